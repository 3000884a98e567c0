/* ecfft_b200 — C ABI of the B200-native ECFFT engine.
 *
 * Drop-in boundary for the hot path of andrewmilson/ecfft: the inherent method surface of
 * `FFTree<secp256k1::Fp>` (reference src/fftree.rs:41-497, re-exported src/lib.rs:10-11),
 * `FftreeField::build_fftree` (src/lib.rs:14-16, 39-85) and the CanonicalSerialize /
 * CanonicalDeserialize layout (src/fftree.rs:510-660).  The reference has no FFI of its own;
 * these are the entry points a Rust shim over `FFTree<Fp>` binds (INTEGRATION.md shows it).
 *
 * Data layout (identical to the reference's in-memory layout so a shim can pass
 * `slice.as_ptr() as *const u64`): a field element is 4 x u64 little-endian limbs in
 * MONTGOMERY form (ark-ff `Fp256<MontBackend<FqConfig,4>>`, src/lib.rs:37), vectors are
 * contiguous arrays of such elements.  Moiety: 0 = S0, 1 = S1 (src/fftree.rs:17-21).
 * Outputs are written to caller-allocated buffers; the library keeps no reference to caller
 * memory after a call returns.
 *
 * Errors: the reference panics ("TODO: errors", src/fftree.rs:40); every entry point here
 * returns a status instead and never aborts.  ecfft_last_error() describes the last failure
 * on the calling thread.
 *
 * `*_dev` variants take DEVICE pointers (16-byte aligned) and a cudaStream_t (as void*; NULL is
 * the legacy default stream, exactly as in the CUDA runtime); they only enqueue work on that
 * stream and do not synchronise.
 */
#ifndef ECFFT_B200_H
#define ECFFT_B200_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ECFFT_OK 0
#define ECFFT_ERR_NOT_POW2 1         /* assert!(n.is_power_of_two())      src/fftree.rs:44,490  */
#define ECFFT_ERR_TREE_TOO_SMALL 2   /* panic!("FFTree is too small")     src/fftree.rs:494     */
#define ECFFT_ERR_BAD_BYTES 3        /* SerializationError                src/fftree.rs:602-660 */
#define ECFFT_ERR_CUDA 4
#define ECFFT_ERR_INVALID_ARG 5
#define ECFFT_ERR_TOO_LARGE 6        /* build_fftree -> None, log2 n >= 36  src/lib.rs:61-64    */
#define ECFFT_ERR_MISSING_TABLES 7   /* handle built with ECFFT_PARTS_ENTER_ONLY               */
#define ECFFT_ERR_BUFFER_TOO_SMALL 8

#define ECFFT_S0 0
#define ECFFT_S1 1

#define ECFFT_PARTS_FULL 0        /* every field of FFTree, as the reference builds it          */
#define ECFFT_PARTS_ENTER_ONLY 1  /* f, matrices, xnn_s(_inv): what ENTER / EXTEND touch        */

typedef struct ecfft_tree ecfft_tree; /* FFTree<Fp> resident in one GPU's HBM, src/fftree.rs:23-38 */

const char* ecfft_last_error(void);
int ecfft_device_count(int* count);

/* ---- construction / persistence ------------------------------------------------------- */
/* <Fp as FftreeField>::build_fftree(n), src/lib.rs:39-85 */
int ecfft_tree_build_secp256k1(size_t n, int parts, int device, ecfft_tree** out);
/* FFTree::new(leaves, rational_maps), src/fftree.rs:42-70.  leaves: n Montgomery elements.
 * Rational map i (src/utils.rs:367-371) is given by map_lens[2i] numerator and map_lens[2i+1]
 * denominator coefficients (low -> high, Montgomery), all concatenated in map_coeffs. */
int ecfft_tree_new(const uint64_t* leaves, size_t n, const uint64_t* map_coeffs, const size_t* map_lens,
                   size_t nmaps, int parts, int device, ecfft_tree** out);
/* CanonicalDeserialize::deserialize_{compressed,uncompressed}, src/fftree.rs:602-660 */
int ecfft_tree_deserialize(const uint8_t* bytes, size_t len, int compressed, int device, ecfft_tree** out);
/* CanonicalSerialize::serialized_size / serialize_with_mode, src/fftree.rs:510-591 */
int ecfft_tree_serialized_size(const ecfft_tree* t, int compressed, size_t* size);
int ecfft_tree_serialize(const ecfft_tree* t, int compressed, uint8_t* buf, size_t cap, size_t* written);
void ecfft_tree_free(ecfft_tree* t);
/* f.leaves().len() of the handle's top tree */
size_t ecfft_tree_leaves(const ecfft_tree* t);
int ecfft_tree_device(const ecfft_tree* t);
/* Read one of the pub fields (src/fftree.rs:25-37) of subtree_with_size(subtree_leaves) as
 * Montgomery elements.  name: "f" (2N), "recombine_matrices" / "decompose_matrices" (4N, row
 * major), "xnn_s", "xnn_s_inv" (N), "z0_s1", "z1_s0", "z0_inv_s1", "z1_inv_s0" (N/2),
 * "z0z0_rem_xnn_s", "z1z1_rem_xnn_s" (N).  out may be NULL to query *count only. */
int ecfft_tree_table(const ecfft_tree* t, size_t subtree_leaves, const char* name, uint64_t* out,
                     size_t cap_elems, size_t* count);

/* ---- the FFTree<F> algorithms, host buffers ------------------------------------------- */
int ecfft_enter(const ecfft_tree* t, const uint64_t* coeffs, size_t n, uint64_t* evals);            /* :164 */
int ecfft_exit(const ecfft_tree* t, const uint64_t* evals, size_t n, uint64_t* coeffs);             /* :227 */
/* `count` calls of ecfft_enter on consecutive vectors (coeffs, evals: count * n elements) as one pipelined call: the
 * upload of vector i+1 and the download of vector i-1 overlap the kernels of vector i.  What a caller's loop around
 * FFTree::enter (src/fftree.rs:164; benches/fftree.rs:28-33 calls it per polynomial) costs on a device. */
int ecfft_enter_many(const ecfft_tree* t, const uint64_t* coeffs, size_t n, size_t count, uint64_t* evals);
int ecfft_extend(const ecfft_tree* t, const uint64_t* evals, size_t n, int moiety, uint64_t* out);  /* :123 */
int ecfft_mextend(const ecfft_tree* t, const uint64_t* evals, size_t n, int moiety, uint64_t* out); /* :138 */
int ecfft_degree(const ecfft_tree* t, const uint64_t* evals, size_t n, size_t* degree);             /* :195 */
/* a has n elements (the reference's zip would silently truncate a longer `a`) */
int ecfft_redc_z0(const ecfft_tree* t, const uint64_t* evals, const uint64_t* a, size_t n, uint64_t* out); /* :264 */
int ecfft_redc_z1(const ecfft_tree* t, const uint64_t* evals, const uint64_t* a, size_t n, uint64_t* out); /* :272 */
int ecfft_modular_reduce(const ecfft_tree* t, const uint64_t* evals, const uint64_t* a, const uint64_t* c,
                         size_t n, uint64_t* out);                                                   /* :286 */
/* out has 2n elements */
int ecfft_vanish(const ecfft_tree* t, const uint64_t* vanish_domain, size_t n, uint64_t* out);      /* :313 */

/* Element-wise product of two vectors of field elements: out[i] = a[i] * b[i] (what `*` on two ark-ff Fp values
 * computes, Montgomery form in and out).  Not a method of FFTree — it is the caller-side step between ENTER
 * and EXIT when polynomials are multiplied through the tree (reference README.md:60-63 usage pattern, the
 * evaluate / interpolate pair of benches/comparison.rs:37-43); it runs on the handle's GPU so the evaluations
 * never leave the device.  Any n. */
int ecfft_pointwise_mul(const ecfft_tree* t, const uint64_t* a, const uint64_t* b, size_t n, uint64_t* out);

/* ---- device-buffer variants (no PCIe traffic; enqueue only) --------------------------- */
int ecfft_enter_dev(const ecfft_tree* t, const void* d_coeffs, size_t n, void* d_evals, void* stream);
int ecfft_exit_dev(const ecfft_tree* t, const void* d_evals, size_t n, void* d_coeffs, void* stream);
int ecfft_extend_dev(const ecfft_tree* t, const void* d_evals, size_t n, int moiety, void* d_out, void* stream);
int ecfft_mextend_dev(const ecfft_tree* t, const void* d_evals, size_t n, int moiety, void* d_out, void* stream);
int ecfft_degree_dev(const ecfft_tree* t, const void* d_evals, size_t n, size_t* degree, void* stream); /* synchronises */
int ecfft_redc_z0_dev(const ecfft_tree* t, const void* d_evals, const void* d_a, size_t n, void* d_out, void* stream);
int ecfft_redc_z1_dev(const ecfft_tree* t, const void* d_evals, const void* d_a, size_t n, void* d_out, void* stream);
int ecfft_modular_reduce_dev(const ecfft_tree* t, const void* d_evals, const void* d_a, const void* d_c,
                             size_t n, void* d_out, void* stream);
int ecfft_vanish_dev(const ecfft_tree* t, const void* d_domain, size_t n, void* d_out, void* stream);
int ecfft_pointwise_mul_dev(const ecfft_tree* t, const void* d_a, const void* d_b, size_t n, void* d_out, void* stream);
/* Multi-GPU building block (DESIGN.md "multi-GPU"): run only the bottom-up ENTER recursion
 * depths whose block size m satisfies m_lo < m <= m_hi on an array of n elements that already
 * holds n/m_lo evaluation vectors of length m_lo (m_lo = 1: raw coefficients); n may be any multiple of
 * m_hi (the blocks are independent).  Rank g runs
 * (1, n/G] on its coefficient chunk, the chunks are all-gathered, then (n/G, n] finishes. */
int ecfft_enter_range_dev(const ecfft_tree* t, const void* d_in, size_t n, size_t m_lo, size_t m_hi,
                          void* d_out, void* stream);

/* Fully sharded top depths (ecfft_b200/dist.py): for the ENTER depth with block size m
 * (EXTEND of length-m/2 vectors towards S1) a rank holds a contiguous chunk of one vector.
 *   prescale : out[e] = in[e] / Gamma^0[pos0 + e]                      (before the decompose levels)
 *   cross    : one butterfly level (phase 0 decompose / 1 recombine, half-stride 2^j) whose pairs
 *              straddle two ranks; role 0: this rank holds the lower elements; p_pos0 = position in
 *              the vector of the first LOWER element; `partner` is the other rank's chunk
 *   local    : all levels with half-stride < count on this rank's chunk
 *   combine  : src/fftree.rs:155-159 for i in [i0, i0+count) of one block, u1/v1 still unscaled;
 *              out has 2*count elements (positions 2*i0 .. of the block) */
int ecfft_mg_prescale_dev(const ecfft_tree* t, size_t m, size_t pos0, const void* d_in, size_t count, void* d_out, void* stream);
int ecfft_mg_cross_dev(const ecfft_tree* t, size_t m, int phase, unsigned j, int role, size_t p_pos0, const void* d_own,
                       const void* d_partner, size_t count, void* d_out, void* stream);
int ecfft_mg_local_dev(const ecfft_tree* t, size_t m, const void* d_in, size_t count, void* d_out, void* stream);
int ecfft_mg_combine_dev(const ecfft_tree* t, size_t m, size_t i0, const void* d_u0, const void* d_v0, const void* d_u1,
                         const void* d_v1, size_t count, void* d_out, void* stream);

/* Peer exchange: instead of a send/recv per straddling level, `d_partner` of
 * ecfft_mg_cross_dev and the u/v pointers of ecfft_mg_combine_dev may point into ANOTHER GPU's memory,
 * mapped through CUDA IPC — the butterfly kernel then loads the partner's operands over NVLink itself.
 *   arena_alloc : zeroed device memory on `device` plus its 64-byte IPC handle (to be sent to the peers)
 *   arena_open  : map a peer's arena into this process (peer access enabled lazily); arena_close unmaps
 *   signal/wait : stream-ordered u64 flags inside arenas (release / acquire at system scope).  A wait not
 *                 satisfied within timeout_ms traps (a CUDA error on the next call, never a hung GPU);
 *                 timeout_ms = 0 waits for ever.  ecfft_enter_peer_dev takes its timeout from the environment
 *                 variable ECFFT_B200_PEER_TIMEOUT_MS (default 20000, 0 = for ever): raise it when ranks can
 *                 reach a call far apart in time (a tree build or data loading on one of them).
 *   arena_reset : after an aborted call (an error on one rank leaves flags and epochs out of step) every rank
 *                 zeroes the flags of its OWN arena, the ranks meet at a host barrier, and epochs restart at 1. */
int ecfft_mg_arena_bytes(size_t n, int world, size_t* bytes);   /* arena size ecfft_enter_peer_dev needs */
int ecfft_mg_arena_alloc(int device, size_t bytes, void** d_ptr, unsigned char* handle64);
int ecfft_mg_arena_open(int device, const unsigned char* handle64, void** d_peer_ptr);
int ecfft_mg_arena_close(void* d_peer_ptr);
int ecfft_mg_arena_free(void* d_ptr);
int ecfft_mg_arena_reset(void* d_ptr, void* stream);
/* 0, or the record of the first wait of ecfft_enter_peer_dev / ecfft_exit_peer_dev on this rank that timed out:
 * bit 63 set, bits 24.. the epoch, bits 8..23 the synchronisation step (0xffff: the wait for the previous call's
 * end), bits 0..7 which peer.  By default such a wait also traps; with ECFFT_B200_PEER_NO_TRAP set it gives up
 * instead and the caller must check this word (synchronises with the device). */
int ecfft_mg_arena_status(const void* d_ptr, unsigned long long* status);
int ecfft_mg_signal_dev(void* d_flag, unsigned long long value, void* stream);
int ecfft_mg_wait_dev(const void* d_flag, unsigned long long value, unsigned timeout_ms, void* stream);
/* The whole per-rank schedule of the sharded ENTER in one call (reference src/fftree.rs:143-161 for the
 * coefficient chunk of `rank`, then the top log2(world) depths over peer memory): d_chunk holds this
 * rank's n/world coefficients, arena_bases[r] is rank r's arena as mapped into this process
 * (arena_bases[rank] = this rank's own), epoch must grow by one per call on all ranks alike;
 * d_out_chunk receives evaluations [rank n/world, (rank+1) n/world).  Calls are ordered against each other
 * on the device: a call first waits until every peer has published the end of the previous one (the last
 * flag of each arena), so no host barrier is needed between calls. */
int ecfft_enter_peer_dev(const ecfft_tree* t, const void* d_chunk, size_t n, int rank, int world, void* const* arena_bases,
                         unsigned long long epoch, void* d_out_chunk, void* stream);

/* EXIT the same way (reference src/fftree.rs:200-224): d_chunk holds evaluations [rank n/world, (rank+1) n/world),
 * d_out_chunk receives the coefficients of the same range.  The top log2(world) depths run MOD (src/fftree.rs:277-281)
 * on vectors spread over the ranks — the butterfly levels whose pairs straddle two ranks read the partner's
 * operands from its arena — then split (u0 | v0), each half going to half of the ranks; below, every rank runs an
 * independent EXIT(n/world).  The arenas must hold ecfft_mg_exit_arena_bytes(n, world); epochs are shared with
 * ecfft_enter_peer_dev (one counter per arena set, growing by one per call of either kind).  Needs a full tree. */
int ecfft_mg_exit_arena_bytes(size_t n, int world, size_t* bytes);
int ecfft_exit_peer_dev(const ecfft_tree* t, const void* d_chunk, size_t n, int rank, int world, void* const* arena_bases,
                        unsigned long long epoch, void* d_out_chunk, void* stream);

/* Device self-test of the field routines the butterflies use for a + b and a - b on unreduced operands
 * (the replacement of ark-ff's add/sub at src/utils.rs:341-346): `samples` directed operand pairs that
 * drive the rare carry paths, checked on the GPU against canonical arithmetic.
 * counters3 = { mismatches, additions that took the rare path, subtractions that did }. */
int ecfft_selftest_field(int device, unsigned long long samples, unsigned long long* counters3);

/* ---- the reference's second field: FFTree<m31::Fp> (src/lib.rs:190-215) ----------------------------------
 * Elements are uint32_t holding the canonical value in [0, 2^31 - 1), which is what the reference's
 * `ark_ff_optimized::fp31::Fp(pub u32)` keeps in memory (src/lib.rs:199-206 writes its constants that way).
 * The handle owns the device tables of every chain level, built on the device by the same sequence as
 * build_ec_fftree / FFTree::new / from_tree (src/ec.rs:498-554, src/fftree.rs:42-70, 318-463).  Same method
 * surface, argument meaning and status codes as the secp256k1 entry points above; ECFFT_ERR_TOO_LARGE when
 * log2 n > 28 (build_ec_fftree returns None, src/ec.rs:513-515). */
typedef struct ecfft_m31_tree ecfft_m31_tree;
int ecfft_m31_tree_build(size_t n, int device, ecfft_m31_tree** out);                                            /* lib.rs:197 */
void ecfft_m31_tree_free(ecfft_m31_tree* t);
size_t ecfft_m31_tree_leaves(const ecfft_m31_tree* t);
/* same table names as ecfft_tree_table; matrices are 4 values each (row major) */
int ecfft_m31_tree_table(const ecfft_m31_tree* t, size_t subtree_leaves, const char* name, uint32_t* out, size_t cap_elems, size_t* count);
int ecfft_m31_enter(const ecfft_m31_tree* t, const uint32_t* coeffs, size_t n, uint32_t* evals);                 /* fftree.rs:164 */
int ecfft_m31_exit(const ecfft_m31_tree* t, const uint32_t* evals, size_t n, uint32_t* coeffs);                  /* :227 */
int ecfft_m31_extend(const ecfft_m31_tree* t, const uint32_t* evals, size_t n, int moiety, uint32_t* out);       /* :123 */
int ecfft_m31_mextend(const ecfft_m31_tree* t, const uint32_t* evals, size_t n, int moiety, uint32_t* out);      /* :138 */
int ecfft_m31_degree(const ecfft_m31_tree* t, const uint32_t* evals, size_t n, size_t* degree);                  /* :195 */
int ecfft_m31_redc_z0(const ecfft_m31_tree* t, const uint32_t* evals, const uint32_t* a, size_t n, uint32_t* out); /* :264 */
int ecfft_m31_redc_z1(const ecfft_m31_tree* t, const uint32_t* evals, const uint32_t* a, size_t n, uint32_t* out); /* :272 */
int ecfft_m31_modular_reduce(const ecfft_m31_tree* t, const uint32_t* evals, const uint32_t* a, const uint32_t* c, size_t n, uint32_t* out); /* :286 */
int ecfft_m31_vanish(const ecfft_m31_tree* t, const uint32_t* domain, size_t n, uint32_t* out);                  /* :313; out has 2n elements */
/* operands resident in HBM, work enqueued on `stream` */
int ecfft_m31_enter_dev(const ecfft_m31_tree* t, const void* d_coeffs, size_t n, void* d_evals, void* stream);
int ecfft_m31_exit_dev(const ecfft_m31_tree* t, const void* d_evals, size_t n, void* d_coeffs, void* stream);
int ecfft_m31_extend_dev(const ecfft_m31_tree* t, const void* d_evals, size_t n, int moiety, void* d_out, void* stream);

/* ---- instrumentation used by bench.py ------------------------------------------------- */
/* kernels launched by this library since it was loaded */
unsigned long long ecfft_launch_count(void);
/* when enabled, every launch of the two hot kernels is bracketed by CUDA events on its stream */
void ecfft_profile_enable(int on);
#define ECFFT_KERNEL_EXTEND_TILE 0
#define ECFFT_KERNEL_ENTER_COMBINE 1
/* sums (and clears) the records of one kernel: device ms, algorithmic bytes, launches */
int ecfft_profile_read(int kernel, double* ms, double* alg_bytes, unsigned long long* launches);

/* Diagnostics of the flow kernel (all passes of an ENTER in one persistent launch): when enabled, its CTAs
 * accumulate {cycles waiting for input blocks, cycles in tile bodies, cycles publishing results, tiles} on the
 * current device.  out4 (may be NULL) receives and clears the accumulators; `enable` switches them on / off. */
int ecfft_flow_stats(int enable, unsigned long long* out4);

#ifdef __cplusplus
}
#endif
#endif
