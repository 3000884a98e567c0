/* ORACLE — TEST INFRASTRUCTURE ONLY.
 *
 * CPU restatement (plain C, gcc) of andrewmilson/ecfft's secp256k1 FFTree path,
 * used by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * `--impl reference` legs as the checker and the timed CPU baseline.  Nothing
 * under ecfft_b200/ may include, link or call this.
 *
 * The Rust reference cannot be built here (no cargo/rustc; arkworks is not
 * vendored), so this file restates it function by function; each function cites
 * the reference file:line it follows.  Third-party arithmetic restated from its
 * published algorithm: ark-ff 0.4 `Fp256<MontBackend<FqConfig,4>>` (4x64-bit
 * CIOS Montgomery, R = 2^256), `batch_inversion` (zeros untouched), `sqrt` for
 * p = 3 mod 4 (x^((p+1)/4), checked); ark-serialize 0.4 wire conventions.
 *
 * Pinning: the reference holds no golden vectors for this path.  The oracle is
 * pinned by re-running the reference's own property tests (src/lib.rs:108-186:
 * ENTER == Horner at every leaf, EXTEND S0<->S1 == Horner, both serialisation
 * round trips) with an independent Python big-integer evaluator
 * (oracle/pyref.py, tests/test_oracle.py), plus the curve constants of
 * src/lib.rs:45-59.  The serialised byte layout itself is "parity unpinned"
 * (no golden bytes exist in the reference).
 */
#ifndef ECFFT_ORACLE_H
#define ECFFT_ORACLE_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Field element in ark-ff memory layout: 4 x u64 little-endian limbs, Montgomery form. */
typedef struct { uint64_t l[4]; } fe;

typedef struct orc_tree orc_tree; /* FFTree<Fp>, reference src/fftree.rs:23-38 */

/* field (ark-ff restatement) */
void orc_fe_mul(const fe* a, const fe* b, fe* r);
void orc_fe_add(const fe* a, const fe* b, fe* r);
void orc_fe_sub(const fe* a, const fe* b, fe* r);
void orc_fe_inv(const fe* a, fe* r);
void orc_fe_from_canonical(const uint8_t bytes_le[32], fe* r); /* canonical int -> Montgomery */
void orc_fe_to_canonical(const fe* a, uint8_t bytes_le[32]);
void orc_batch_inversion(fe* v, size_t n);

/* Fp::build_fftree(n), reference src/lib.rs:39-85.  NULL when log2 n >= 36.
 * parts: 0 = full tree (reference behaviour); 1 = only what ENTER/EXTEND(S1)
 * touch (f, matrices, xnn_s, xnn_s_inv) — used for the large CPU-baseline trees. */
orc_tree* orc_build_fftree(size_t n, int parts);
/* threads the element-wise loops of the tree build may use (default 1); tables are identical for any value */
void orc_set_build_threads(int threads);
void orc_tree_free(orc_tree* t);
size_t orc_tree_leaves(const orc_tree* t);
const orc_tree* orc_subtree_with_size(const orc_tree* t, size_t n); /* src/fftree.rs:489-496; NULL if too small */

/* table access (Montgomery limbs), name in {"f","recombine","decompose","xnn_s","xnn_s_inv",
 * "z0_s1","z1_s0","z0_inv_s1","z1_inv_s0","z0z0_rem_xnn_s","z1z1_rem_xnn_s"}; returns element count
 * (matrices count 4 elements each) and sets *ptr. */
size_t orc_tree_table(const orc_tree* t, const char* name, const fe** ptr);

/* the eight algorithms, reference src/fftree.rs:72-316.  moiety: 0 = S0, 1 = S1.
 * Return 0 on success, nonzero where the reference panics. */
int orc_enter(const orc_tree* t, const fe* coeffs, size_t n, fe* out);
int orc_exit(const orc_tree* t, const fe* evals, size_t n, fe* out);
int orc_extend(const orc_tree* t, const fe* evals, size_t n, int moiety, fe* out);
int orc_mextend(const orc_tree* t, const fe* evals, size_t n, int moiety, fe* out);
int orc_degree(const orc_tree* t, const fe* evals, size_t n, size_t* degree);
int orc_redc_z0(const orc_tree* t, const fe* evals, const fe* a, size_t n, fe* out);
int orc_redc_z1(const orc_tree* t, const fe* evals, const fe* a, size_t n, fe* out);
int orc_modular_reduce(const orc_tree* t, const fe* evals, const fe* a, const fe* c, size_t n, fe* out);
int orc_vanish(const orc_tree* t, const fe* domain, size_t n, fe* out /* 2n */);
/* ENTER with the two independent half-recursions / EXTENDs forked on up to `threads` pthreads
 * (same arithmetic, same result; used as the all-cores CPU baseline). */
int orc_enter_mt(const orc_tree* t, const fe* coeffs, size_t n, fe* out, int threads);

/* bottom-up ENTER restricted to recursion depths with block size m_lo < m <= m_hi (schedule check) */
int orc_enter_range(const orc_tree* t, const fe* in, size_t n, size_t m_lo, size_t m_hi, fe* out);

/* CanonicalSerialize / CanonicalDeserialize, reference src/fftree.rs:510-660 */
size_t orc_serialized_size(const orc_tree* t, int compressed);
size_t orc_serialize(const orc_tree* t, int compressed, uint8_t* buf, size_t cap);
orc_tree* orc_deserialize(const uint8_t* buf, size_t len, int compressed);

#ifdef __cplusplus
}
#endif
#endif
