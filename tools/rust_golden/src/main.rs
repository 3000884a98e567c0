//! Golden vectors from the real reference crate (andrewmilson/ecfft @ 9ca932a, arkworks 0.4) for
//! `FFTree<secp256k1::Fp>` at n = 64 with `StdRng::from_seed([1; 32])` — the tree size and seed of the
//! reference's own tests (src/lib.rs:102-186).  NOT BUILT HERE (no Rust toolchain in this image); see
//! Cargo.toml for how to run it.  Output (stdout) is the line format tests/test_arkworks_golden.py reads:
//!
//!   n <leaves>
//!   bytes <name> <hex>            serialised FFTree (CanonicalSerialize, src/fftree.rs:510-554)
//!   vec <name> <hex>              Vec<Fp>: 32 bytes per element = the 4 u64 limbs, little endian, exactly as
//!                                 ark-ff holds them in memory (Montgomery form, src/lib.rs:37)
//!   num <name> <decimal>
use ark_ff::UniformRand;
use ark_serialize::CanonicalSerialize;
use ecfft::secp256k1::Fp;
use ecfft::FftreeField;
use ecfft::Moiety;
use rand::rngs::StdRng;
use rand::SeedableRng;

fn hex(bytes: &[u8]) -> String {
    bytes.iter().map(|b| format!("{:02x}", b)).collect()
}

fn limbs(v: &[Fp]) -> String {
    let mut out = Vec::with_capacity(v.len() * 32);
    for x in v {
        // Fp<MontBackend<_, 4>, 4>(BigInt<4>([u64; 4]), PhantomData): the in-memory Montgomery limbs
        for limb in (x.0).0.iter() {
            out.extend_from_slice(&limb.to_le_bytes());
        }
    }
    hex(&out)
}

fn vec_line(name: &str, v: &[Fp]) {
    println!("vec {} {}", name, limbs(v));
}

/// The same for the second field (src/lib.rs:190-215): `vec32 m31.<name> <hex>`, 4 bytes per element = the u32 that
/// `ark_ff_optimized::fp31::Fp(pub u32)` holds.  tests/test_m31.py reads these lines from tests/golden/arkworks_n64.txt (this
/// program's whole output) or from tests/golden/arkworks_m31_n64.txt.
fn m31_vectors() {
    use ecfft::m31::Fp as F31;
    let n = 64usize;
    let fftree = F31::build_fftree(n).unwrap();
    let mut rng = StdRng::from_seed([1; 32]);
    let v32 = |name: &str, v: &[F31]| {
        let mut out = Vec::with_capacity(v.len() * 4);
        for x in v {
            out.extend_from_slice(&x.0.to_le_bytes());
        }
        println!("vec32 m31.{} {}", name, hex(&out));
    };
    v32("leaves", &fftree.subtree_with_size(n).eval_domain());
    v32("xnn_s", &fftree.xnn_s);
    v32("z0z0_rem_xnn_s", &fftree.z0z0_rem_xnn_s);
    let coeffs: Vec<F31> = (0..n).map(|_| F31::rand(&mut rng)).collect();
    let evals = fftree.enter(&coeffs);
    v32("enter.in", &coeffs);
    v32("enter.out", &evals);
    let arbitrary: Vec<F31> = (0..n).map(|_| F31::rand(&mut rng)).collect();
    v32("exit.in", &arbitrary);
    v32("exit.out", &fftree.exit(&arbitrary));
    let half: Vec<F31> = (0..n / 2).map(|_| F31::rand(&mut rng)).collect();
    v32("extend.in", &half);
    v32("extend_s1.out", &fftree.extend(&half, Moiety::S1));
    v32("extend_s0.out", &fftree.extend(&half, Moiety::S0));
    v32("mextend_s1.out", &fftree.mextend(&half, Moiety::S1));
    v32("redc_z0.out", &fftree.redc_z0(&arbitrary, &fftree.xnn_s));
    v32("mod.out", &fftree.modular_reduce(&arbitrary, &fftree.xnn_s, &fftree.z0z0_rem_xnn_s));
    v32("vanish.out", &fftree.vanish(&half));
    println!("num m31.degree.out {}", fftree.degree(&evals));
}

fn main() {
    m31_vectors();
    let n = 64usize;
    let fftree = Fp::build_fftree(n).unwrap();
    let mut rng = StdRng::from_seed([1; 32]);
    println!("n {}", n);

    let mut compressed = Vec::new();
    fftree.serialize_compressed(&mut compressed).unwrap();
    println!("bytes tree_compressed {}", hex(&compressed));
    let mut uncompressed = Vec::new();
    fftree.serialize_uncompressed(&mut uncompressed).unwrap();
    println!("bytes tree_uncompressed {}", hex(&uncompressed));

    vec_line("leaves", &fftree.subtree_with_size(n).eval_domain());
    vec_line("xnn_s", &fftree.xnn_s);
    vec_line("z0z0_rem_xnn_s", &fftree.z0z0_rem_xnn_s);

    // ENTER / EXIT (src/fftree.rs:164, 227)
    let coeffs: Vec<Fp> = (0..n).map(|_| Fp::rand(&mut rng)).collect();
    let evals = fftree.enter(&coeffs);
    vec_line("enter.in", &coeffs);
    vec_line("enter.out", &evals);
    let arbitrary: Vec<Fp> = (0..n).map(|_| Fp::rand(&mut rng)).collect();
    vec_line("exit.in", &arbitrary);
    vec_line("exit.out", &fftree.exit(&arbitrary));

    // EXTEND / MEXTEND both ways on the 64-leaf tree (inputs of n / 2 values, src/fftree.rs:123, 138)
    let half: Vec<Fp> = (0..n / 2).map(|_| Fp::rand(&mut rng)).collect();
    vec_line("extend.in", &half);
    vec_line("extend_s1.out", &fftree.extend(&half, Moiety::S1));
    vec_line("extend_s0.out", &fftree.extend(&half, Moiety::S0));
    vec_line("mextend_s1.out", &fftree.mextend(&half, Moiety::S1));
    vec_line("mextend_s0.out", &fftree.mextend(&half, Moiety::S0));

    // REDC / MOD with the tree's own tables as in benches/fftree.rs:48-54, but with matching lengths
    vec_line("redc.in", &arbitrary);
    vec_line("redc_z0.out", &fftree.redc_z0(&arbitrary, &fftree.xnn_s));
    vec_line("redc_z1.out", &fftree.redc_z1(&arbitrary, &fftree.xnn_s));
    vec_line("mod.out", &fftree.modular_reduce(&arbitrary, &fftree.xnn_s, &fftree.z0z0_rem_xnn_s));
    // and with arbitrary `a`, `c`
    let a: Vec<Fp> = (0..n).map(|_| Fp::rand(&mut rng)).collect();
    let c: Vec<Fp> = (0..n).map(|_| Fp::rand(&mut rng)).collect();
    vec_line("redc_a.in", &a);
    vec_line("mod_c.in", &c);
    vec_line("redc_z0_a.out", &fftree.redc_z0(&arbitrary, &a));
    vec_line("mod_ac.out", &fftree.modular_reduce(&arbitrary, &a, &c));

    // VANISH of n / 2 points on the 64-leaf tree (src/fftree.rs:313) and DEGREE (src/fftree.rs:195)
    vec_line("vanish.in", &half);
    vec_line("vanish.out", &fftree.vanish(&half));
    let mut low = coeffs.clone();
    for x in low.iter_mut().skip(41) {
        *x = Fp::from(0u64);
    }
    let low_evals = fftree.enter(&low);
    vec_line("degree.in", &low_evals);
    println!("num degree.out {}", fftree.degree(&low_evals));
    println!("num degree_full.out {}", fftree.degree(&evals));
}
