// extern "C" boundary (include/ecfft_b200.h).  Maps engine exceptions to status codes and moves
// host buffers through the handle's stream.  No torch types, no CPU fallback: every entry point
// that computes needs a CUDA device and fails with ECFFT_ERR_CUDA without one.
#include <string.h>
#include <stdio.h>
#include <stdlib.h>

#include <chrono>

#include <mutex>

#include "../../include/ecfft_b200.h"
#include "engine.h"

using namespace ecfft;

struct ecfft_tree {
  Tree* tree;
  std::mutex mu;  // host-buffer calls share the handle's stream and are serialised
};

static thread_local std::string g_last_error;
namespace ecfft { void set_last_error(const char* msg) { g_last_error = msg ? msg : ""; } }   // for the other ABI files (m31.cu)

template <class F>
static int guard(F f) {
  try {
    f();
    return ECFFT_OK;
  } catch (const Error& e) {
    g_last_error = e.what();
    return e.code;
  } catch (const std::exception& e) {
    g_last_error = e.what();
    return ECFFT_ERR_INVALID_ARG;
  }
}
static void require(bool ok, int code, const char* msg) {
  if (!ok) throw Error(code, msg);
}
static const Fp* dptr(const void* p) {
  require(p != nullptr && ((uintptr_t)p & 15) == 0, ERR_INVALID_ARG, "device pointer must be non-null and 16-byte aligned");
  return (const Fp*)p;
}
static Fp* dptr(void* p) { return (Fp*)dptr((const void*)p); }
// `_dev` entry points enqueue on exactly the stream they are given (NULL = the legacy default stream,
// which is what torch's default stream is), so the caller's events and ordering apply.
static cudaStream_t pick_stream(const ecfft_tree*, void* stream) { return (cudaStream_t)stream; }

namespace {
// scoped device buffer fed from / drained to host memory on the handle's stream
struct HostIO {
  DeviceGuard dev;
  const Tree& t;
  cudaStream_t st;
  std::vector<void*> bufs;
  HostIO(const Tree& tree) : dev(tree.device), t(tree), st(tree.stream) {}
  ~HostIO() {
    for (void* b : bufs) cudaFreeAsync(b, st);
    cudaStreamSynchronize(st);
  }
  Fp* alloc(size_t count) {
    void* p = nullptr;
    ECFFT_CUDA(cudaMallocAsync(&p, (count ? count : 1) * sizeof(Fp), st));
    bufs.push_back(p);
    return (Fp*)p;
  }
  Fp* in(const uint64_t* host, size_t count) {
    require(host != nullptr || count == 0, ERR_INVALID_ARG, "null input buffer");
    Fp* d = alloc(count);
    if (count) ECFFT_CUDA(cudaMemcpyAsync(d, host, count * sizeof(Fp), cudaMemcpyHostToDevice, st));
    return d;
  }
  void out(uint64_t* host, const Fp* d, size_t count) {
    require(host != nullptr || count == 0, ERR_INVALID_ARG, "null output buffer");
    if (count) ECFFT_CUDA(cudaMemcpyAsync(host, d, count * sizeof(Fp), cudaMemcpyDeviceToHost, st));
    ECFFT_CUDA(cudaStreamSynchronize(st));
  }
};
}  // namespace

extern "C" {

const char* ecfft_last_error(void) { return g_last_error.c_str(); }

unsigned long long ecfft_launch_count(void) { return prof::launches(); }
void ecfft_profile_enable(int on) { prof::enable(on != 0); }
int ecfft_profile_read(int kernel, double* ms, double* alg_bytes, unsigned long long* launches) {
  return guard([&] {
    require(kernel >= 0 && kernel < prof::NUM_KERNELS && ms && alg_bytes && launches, ERR_INVALID_ARG, "bad profile query");
    prof::read((prof::Kernel)kernel, ms, alg_bytes, launches);
  });
}

int ecfft_flow_stats(int enable, unsigned long long* out4) {
  return guard([&] {
    if (out4) k::flow_stats_read(out4);
    k::flow_stats_enable(enable != 0);
  });
}

int ecfft_device_count(int* count) {
  return guard([&] {
    require(count != nullptr, ERR_INVALID_ARG, "null count");
    ECFFT_CUDA(cudaGetDeviceCount(count));
  });
}

int ecfft_tree_build_secp256k1(size_t n, int parts, int device, ecfft_tree** out) {
  return guard([&] {
    require(out != nullptr, ERR_INVALID_ARG, "null out");
    require(parts == PARTS_FULL || parts == PARTS_ENTER_ONLY, ERR_INVALID_ARG, "bad parts");
    // argument errors come before anything that needs a device (src/fftree.rs:44, src/lib.rs:61-64)
    require(n && !(n & (n - 1)), ERR_NOT_POW2, "n is not a power of two");
    require(n < ((size_t)1 << 36), ERR_TOO_LARGE, "FFTree size is too large for the generator (log2 n >= 36)");
    DeviceGuard dev_guard(device);
    ecfft_tree* h = new ecfft_tree();
    try {
      h->tree = build_secp256k1(n, parts, device);
    } catch (...) {
      delete h;
      throw;
    }
    *out = h;
  });
}

int ecfft_tree_new(const uint64_t* leaves, size_t n, const uint64_t* map_coeffs, const size_t* map_lens, size_t nmaps,
                   int parts, int device, ecfft_tree** out) {
  return guard([&] {
    require(out && leaves && (nmaps == 0 || (map_coeffs && map_lens)), ERR_INVALID_ARG, "null argument");
    require(n && !(n & (n - 1)), ERR_NOT_POW2, "leaf count is not a power of two");
    require(parts == PARTS_FULL || parts == PARTS_ENTER_ONLY, ERR_INVALID_ARG, "bad parts");
    DeviceGuard dev_guard(device);
    // Montgomery -> plain: leaves on the device, the few map coefficients on the host
    std::vector<RatMapHost> maps(nmaps);
    const Fp rinv = fp_const_RINV();
    size_t off = 0;
    for (size_t i = 0; i < nmaps; i++)
      for (int part = 0; part < 2; part++) {
        size_t cnt = map_lens[2 * i + part];
        std::vector<Fp>& dst = part == 0 ? maps[i].num : maps[i].den;
        dst.resize(cnt);
        for (size_t c = 0; c < cnt; c++) {
          Fp x;
          memcpy(&x, map_coeffs + 4 * (off + c), sizeof(Fp));
          require(fp_eq(x, fp_canon(x)), ERR_INVALID_ARG, "rational-map coefficient is not a canonical field element");
          dst[c] = fp_mul(x, rinv);
        }
        off += cnt;
      }
    Fp* d = nullptr;
    ECFFT_CUDA(cudaMalloc((void**)&d, n * sizeof(Fp)));
    ecfft_tree* h = new ecfft_tree();
    try {
      ECFFT_CUDA(cudaMemcpy(d, leaves, n * sizeof(Fp), cudaMemcpyHostToDevice));
      {  // ark-ff elements are canonical (< p); anything else is not a field element
        unsigned long long* cnt = nullptr;
        unsigned long long bad = 0;
        ECFFT_CUDA(cudaMalloc((void**)&cnt, sizeof bad));
        cudaError_t e = cudaMemset(cnt, 0, sizeof bad);
        if (e == cudaSuccess) {
          k::count_noncanonical(cnt, d, n, nullptr);
          e = cudaMemcpy(&bad, cnt, sizeof bad, cudaMemcpyDeviceToHost);
        }
        cudaFree(cnt);
        ECFFT_CUDA(e);
        require(bad == 0, ERR_INVALID_ARG, "leaves are not canonical field elements (limbs >= p)");
      }
      k::mul_const(d, d, rinv, n, nullptr);
      ECFFT_CUDA(cudaDeviceSynchronize());
      h->tree = tree_from_leaves(d, n, maps, parts, device);
    } catch (...) {
      cudaFree(d);
      delete h;
      throw;
    }
    cudaFree(d);
    *out = h;
  });
}

int ecfft_tree_deserialize(const uint8_t* bytes, size_t len, int compressed, int device, ecfft_tree** out) {
  return guard([&] {
    require(out && bytes, ERR_INVALID_ARG, "null argument");
    DeviceGuard dev_guard(device);
    ecfft_tree* h = new ecfft_tree();
    try {
      h->tree = deserialize(bytes, len, compressed != 0, device);
    } catch (...) {
      delete h;
      throw;
    }
    *out = h;
  });
}

int ecfft_tree_serialized_size(const ecfft_tree* t, int compressed, size_t* size) {
  return guard([&] {
    require(t && size, ERR_INVALID_ARG, "null argument");
    *size = serialized_size(*t->tree, compressed != 0);
  });
}

int ecfft_tree_serialize(const ecfft_tree* t, int compressed, uint8_t* buf, size_t cap, size_t* written) {
  return guard([&] {
    require(t && buf && written, ERR_INVALID_ARG, "null argument");
    std::lock_guard<std::mutex> lock(const_cast<ecfft_tree*>(t)->mu);
    DeviceGuard dev_guard(t->tree->device);
    *written = serialize(*t->tree, compressed != 0, buf, cap);
  });
}

void ecfft_tree_free(ecfft_tree* t) {
  if (!t) return;
  delete t->tree;
  delete t;
}

size_t ecfft_tree_leaves(const ecfft_tree* t) { return t ? t->tree->n() : 0; }
int ecfft_tree_device(const ecfft_tree* t) { return t ? t->tree->device : -1; }

int ecfft_tree_table(const ecfft_tree* t, size_t subtree_leaves, const char* name, uint64_t* out, size_t cap_elems, size_t* count) {
  return guard([&] {
    require(t && name && count, ERR_INVALID_ARG, "null argument");
    std::lock_guard<std::mutex> lock(const_cast<ecfft_tree*>(t)->mu);
    const Tree& tr = *t->tree;
    HostIO io(tr);
    Engine eng(tr, io.st);
    const Level& lv = eng.level_for(subtree_leaves);
    const size_t N = subtree_leaves, h = N / 2, n = tr.n();
    const Fp* src = nullptr;
    size_t cnt = 0;
    Fp* staged = nullptr;
    std::string nm(name);
    if (nm == "f") {
      cnt = 2 * N;
      if (out) {
        staged = io.alloc(cnt);
        ECFFT_CUDA(cudaMemsetAsync(staged, 0, sizeof(Fp), io.st));
        for (uint32_t kk = 0; kk <= lv.log_n; kk++) k::copy_strided(staged + (N >> kk), tr.f + (n >> kk), N >> kk, n / N, io.st);
        src = staged;
      }
    } else if (nm == "recombine_matrices") { src = lv.rmat; cnt = 4 * N; }
    else if (nm == "decompose_matrices") { src = lv.dmat; cnt = 4 * N; }
    else if (nm == "xnn_s") { src = lv.xnn_s; cnt = N; }
    else if (nm == "xnn_s_inv") { src = lv.xnn_s_inv; cnt = N; }
    else if (nm == "z0_s1") { src = lv.z0_s1; cnt = h; }
    else if (nm == "z1_s0") { src = lv.z1_s0; cnt = h; }
    else if (nm == "z0_inv_s1") { src = lv.z0_inv_s1; cnt = h; }
    else if (nm == "z1_inv_s0") { src = lv.z1_inv_s0; cnt = h; }
    else if (nm == "z0z0_rem_xnn_s") { src = lv.z0z0; cnt = N > 1 ? N : 0; }
    else if (nm == "z1z1_rem_xnn_s") { src = lv.z1z1; cnt = N > 1 ? N : 0; }
    else throw Error(ERR_INVALID_ARG, "unknown table name");
    if (nm != "f" && cnt && !src) throw Error(ERR_MISSING_TABLES, "table was not built");
    *count = cnt;
    if (!out) return;
    require(cap_elems >= cnt, ERR_BUFFER_TOO_SMALL, "table buffer too small");
    Fp* m = io.alloc(cnt);
    k::mul_const(m, src, fp_const_R(), cnt, io.st);  // plain -> Montgomery, the reference's in-memory form
    io.out(out, m, cnt);
  });
}

// ---- host-buffer algorithms -----------------------------------------------------------------
#define LOCKED_IO                                              \
  require(t != nullptr, ERR_INVALID_ARG, "null tree handle");  \
  std::lock_guard<std::mutex> lock(const_cast<ecfft_tree*>(t)->mu); \
  HostIO io(*t->tree);                                         \
  Engine eng(*t->tree, io.st);

int ecfft_enter(const ecfft_tree* t, const uint64_t* coeffs, size_t n, uint64_t* evals) {
  return guard([&] {
    LOCKED_IO
    eng.level_for(n);
    require((coeffs != nullptr && evals != nullptr) || n == 0, ERR_INVALID_ARG, "null buffer");
    Fp* d_out = io.alloc(n);
    // Large inputs: the coefficient vector is uploaded in three chunks (3/16, 5/16, 8/16 of it) on a second
    // stream and the recursion depths with block size <= 1024 (independent contiguous blocks, one launch per
    // depth) run on each chunk while the next is still in flight — the chunks grow at the ratio of the
    // low-depth compute rate to the PCIe rate, so every upload but the first hides behind the previous
    // chunk's work; the deeper depths run once on the whole vector.  Exposed: the first 3/16 of the upload and
    // the final download.  Splitting the deeper depths too costs more in extra launches on an under-filled GPU
    // than it hides (measured at n = 2^22: whole-vector ENTER 15.5 ms, two full half-ENTERs + merge 16.2 ms,
    // n/8,n/8,n/4,n/2 17.6 ms; tools/pcie_probe.py).
    static const bool host_pipe = !(getenv("ECFFT_B200_HOST_PIPE") && atoi(getenv("ECFFT_B200_HOST_PIPE")) == 0);
    static const bool host_trace = getenv("ECFFT_B200_HOST_TRACE") != nullptr;   // prints the call's device timeline (debugging aid)
    if (n >= ((size_t)1 << 16) && host_pipe) {
      const size_t m_split = 1024;
      const size_t off[4] = {0, 3 * (n / 16), 8 * (n / 16), n};
      Fp* d_in = io.alloc(n);
      Fp* d_mid = io.alloc(n);
      cudaStream_t cs = nullptr;
      ECFFT_CUDA(cudaStreamCreateWithFlags(&cs, cudaStreamNonBlocking));
      cudaEvent_t ev[3] = {nullptr, nullptr, nullptr};
      cudaEvent_t tr[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
      const std::chrono::steady_clock::time_point t_host0 = std::chrono::steady_clock::now();
      try {
        if (host_trace) {
          for (int i = 0; i < 8; i++) ECFFT_CUDA(cudaEventCreate(&tr[i]));
          ECFFT_CUDA(cudaEventRecord(tr[0], io.st));
          ECFFT_CUDA(cudaStreamWaitEvent(cs, tr[0], 0));
        }
        for (int g = 0; g < 3; g++) {
          ECFFT_CUDA(cudaEventCreateWithFlags(&ev[g], cudaEventDisableTiming));
          ECFFT_CUDA(cudaMemcpyAsync(d_in + off[g], coeffs + 4 * off[g], (off[g + 1] - off[g]) * sizeof(Fp), cudaMemcpyHostToDevice, cs));
          ECFFT_CUDA(cudaEventRecord(ev[g], cs));
          if (host_trace) ECFFT_CUDA(cudaEventRecord(tr[1 + g], cs));
        }
        for (int g = 0; g < 3; g++) {
          ECFFT_CUDA(cudaStreamWaitEvent(io.st, ev[g], 0));
          eng.enter_range(d_in + off[g], d_mid + off[g], off[g + 1] - off[g], 1, m_split);
          if (host_trace) ECFFT_CUDA(cudaEventRecord(tr[4 + g], io.st));
        }
        eng.enter_range(d_mid, d_out, n, m_split, n);
        if (host_trace) ECFFT_CUDA(cudaEventRecord(tr[7], io.st));
        const double ms_enqueue = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_host0).count();
        io.out(evals, d_out, n);
        if (host_trace) {
          float t[8] = {0};
          for (int i = 1; i < 8; i++) cudaEventElapsedTime(&t[i], tr[0], tr[i]);
          const double ms_total = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_host0).count();
          fprintf(stderr, "[ecfft_enter trace] uploads done %.2f %.2f %.2f | chunk depths done %.2f %.2f %.2f | deep depths done %.2f | host: enqueued %.2f, returned %.2f ms\n",
                  t[1], t[2], t[3], t[4], t[5], t[6], t[7], ms_enqueue, ms_total);
        }
      } catch (...) {
        cudaStreamSynchronize(cs);
        cudaStreamDestroy(cs);
        for (int g = 0; g < 3; g++) if (ev[g]) cudaEventDestroy(ev[g]);
        for (int i = 0; i < 8; i++) if (tr[i]) cudaEventDestroy(tr[i]);
        throw;
      }
      for (int g = 0; g < 3; g++) cudaEventDestroy(ev[g]);
      for (int i = 0; i < 8; i++) if (tr[i]) cudaEventDestroy(tr[i]);
      cudaStreamDestroy(cs);
      return;
    }
    Fp* d_in = io.in(coeffs, n);
    eng.enter(d_in, d_out, n);
    io.out(evals, d_out, n);
  });
}
// `count` coefficient vectors of n elements each, contiguous in host memory, to `count` evaluation vectors: the
// loop a caller of FFTree::enter (src/fftree.rs:164) writes around it, as ONE call, so that the upload of vector
// i+1 and the download of vector i-1 run on the copy engines while vector i is in the kernels (PCIe is full
// duplex): per vector the host-to-host cost tends to the device time instead of device time + both copies.
int ecfft_enter_many(const ecfft_tree* t, const uint64_t* coeffs, size_t n, size_t count, uint64_t* evals) {
  return guard([&] {
    LOCKED_IO
    eng.level_for(n);
    require((coeffs != nullptr && evals != nullptr) || n == 0 || count == 0, ERR_INVALID_ARG, "null buffer");
    if (count == 0 || n == 0) return;
    const int S = count > 1 ? 2 : 1;
    Fp* d_in[2] = {io.alloc(n), S > 1 ? io.alloc(n) : nullptr};
    Fp* d_out[2] = {io.alloc(n), S > 1 ? io.alloc(n) : nullptr};
    cudaStream_t up = nullptr, down = nullptr;
    cudaEvent_t u_done[2] = {nullptr, nullptr}, c_done[2] = {nullptr, nullptr}, d_done[2] = {nullptr, nullptr};
    auto cleanup = [&] {
      if (up) { cudaStreamSynchronize(up); cudaStreamDestroy(up); }
      if (down) { cudaStreamSynchronize(down); cudaStreamDestroy(down); }
      for (int i = 0; i < 2; i++) {
        if (u_done[i]) cudaEventDestroy(u_done[i]);
        if (c_done[i]) cudaEventDestroy(c_done[i]);
        if (d_done[i]) cudaEventDestroy(d_done[i]);
      }
    };
    try {
      ECFFT_CUDA(cudaStreamCreateWithFlags(&up, cudaStreamNonBlocking));
      ECFFT_CUDA(cudaStreamCreateWithFlags(&down, cudaStreamNonBlocking));
      for (int i = 0; i < 2; i++) {
        ECFFT_CUDA(cudaEventCreateWithFlags(&u_done[i], cudaEventDisableTiming));
        ECFFT_CUDA(cudaEventCreateWithFlags(&c_done[i], cudaEventDisableTiming));
        ECFFT_CUDA(cudaEventCreateWithFlags(&d_done[i], cudaEventDisableTiming));
      }
      // the buffers were allocated on io.st: the copy streams start after that point
      ECFFT_CUDA(cudaEventRecord(c_done[0], io.st));
      ECFFT_CUDA(cudaStreamWaitEvent(up, c_done[0], 0));
      ECFFT_CUDA(cudaStreamWaitEvent(down, c_done[0], 0));
      for (size_t i = 0; i < count; i++) {
        const int b = (int)(i % (size_t)S);
        if (i >= (size_t)S) ECFFT_CUDA(cudaStreamWaitEvent(up, c_done[b], 0));        // vector i-S has left d_in[b]
        ECFFT_CUDA(cudaMemcpyAsync(d_in[b], coeffs + 4 * i * n, n * sizeof(Fp), cudaMemcpyHostToDevice, up));
        ECFFT_CUDA(cudaEventRecord(u_done[b], up));
        ECFFT_CUDA(cudaStreamWaitEvent(io.st, u_done[b], 0));
        if (i >= (size_t)S) ECFFT_CUDA(cudaStreamWaitEvent(io.st, d_done[b], 0));     // vector i-S has left d_out[b]
        eng.enter(d_in[b], d_out[b], n);
        ECFFT_CUDA(cudaEventRecord(c_done[b], io.st));
        ECFFT_CUDA(cudaStreamWaitEvent(down, c_done[b], 0));
        ECFFT_CUDA(cudaMemcpyAsync(evals + 4 * i * n, d_out[b], n * sizeof(Fp), cudaMemcpyDeviceToHost, down));
        ECFFT_CUDA(cudaEventRecord(d_done[b], down));
      }
      ECFFT_CUDA(cudaStreamSynchronize(down));
      ECFFT_CUDA(cudaStreamSynchronize(io.st));
    } catch (...) {
      cleanup();
      throw;
    }
    cleanup();
  });
}
int ecfft_exit(const ecfft_tree* t, const uint64_t* evals, size_t n, uint64_t* coeffs) {
  return guard([&] {
    LOCKED_IO
    eng.level_for(n);
    Fp* d_in = io.in(evals, n);
    Fp* d_out = io.alloc(n);
    eng.exit(d_in, d_out, n);
    io.out(coeffs, d_out, n);
  });
}
int ecfft_extend(const ecfft_tree* t, const uint64_t* evals, size_t n, int moiety, uint64_t* out) {
  return guard([&] {
    LOCKED_IO
    require(moiety == 0 || moiety == 1, ERR_INVALID_ARG, "bad moiety");
    require(n > 0 && n <= ((size_t)1 << 62), ERR_NOT_POW2, "bad length");
    eng.level_for(2 * n);
    Fp* d = io.in(evals, n);
    eng.extend(d, d, n, 1, (Moiety)moiety);
    io.out(out, d, n);
  });
}
int ecfft_mextend(const ecfft_tree* t, const uint64_t* evals, size_t n, int moiety, uint64_t* out) {
  return guard([&] {
    LOCKED_IO
    require(moiety == 0 || moiety == 1, ERR_INVALID_ARG, "bad moiety");
    require(n > 0 && n <= ((size_t)1 << 62), ERR_NOT_POW2, "bad length");
    eng.level_for(2 * n);
    Fp* d = io.in(evals, n);
    eng.mextend(d, d, n, (Moiety)moiety, FORM_MONT);
    io.out(out, d, n);
  });
}
int ecfft_degree(const ecfft_tree* t, const uint64_t* evals, size_t n, size_t* degree) {
  return guard([&] {
    LOCKED_IO
    require(degree != nullptr, ERR_INVALID_ARG, "null degree");
    eng.level_for(n);
    Fp* d = io.in(evals, n);
    *degree = eng.degree(d, n);
  });
}
static int redc_host(const ecfft_tree* t, const uint64_t* evals, const uint64_t* a, size_t n, uint64_t* out, Moiety m) {
  return guard([&] {
    LOCKED_IO
    eng.level_for(n);
    Fp* d = io.in(evals, n);
    Fp* da = io.in(a, n);
    Fp* o = io.alloc(n);
    eng.redc_user(d, da, n, m, o);
    io.out(out, o, n);
  });
}
int ecfft_redc_z0(const ecfft_tree* t, const uint64_t* evals, const uint64_t* a, size_t n, uint64_t* out) { return redc_host(t, evals, a, n, out, S0); }
int ecfft_redc_z1(const ecfft_tree* t, const uint64_t* evals, const uint64_t* a, size_t n, uint64_t* out) { return redc_host(t, evals, a, n, out, S1); }
int ecfft_modular_reduce(const ecfft_tree* t, const uint64_t* evals, const uint64_t* a, const uint64_t* c, size_t n, uint64_t* out) {
  return guard([&] {
    LOCKED_IO
    eng.level_for(n);
    Fp* d = io.in(evals, n);
    Fp* da = io.in(a, n);
    Fp* dc = io.in(c, n);
    Fp* o = io.alloc(n);
    eng.mod_user(d, da, dc, n, o);
    io.out(out, o, n);
  });
}
int ecfft_pointwise_mul(const ecfft_tree* t, const uint64_t* a, const uint64_t* b, size_t n, uint64_t* out) {
  return guard([&] {
    LOCKED_IO
    Fp* da = io.in(a, n);
    Fp* db = io.in(b, n);
    k::mul_mont(da, da, db, n, io.st);
    io.out(out, da, n);
  });
}
int ecfft_vanish(const ecfft_tree* t, const uint64_t* vanish_domain, size_t n, uint64_t* out) {
  return guard([&] {
    LOCKED_IO
    require(n > 0 && n <= ((size_t)1 << 62), ERR_NOT_POW2, "bad length");
    eng.level_for(2 * n);
    Fp* d = io.in(vanish_domain, n);
    Fp* o = io.alloc(2 * n);
    eng.vanish(d, o, n, FORM_MONT);
    io.out(out, o, 2 * n);
  });
}

// ---- device-buffer algorithms -----------------------------------------------------------------
#define DEV_ENGINE                                             \
  require(t != nullptr, ERR_INVALID_ARG, "null tree handle");  \
  DeviceGuard dev_guard(t->tree->device);                      \
  Engine eng(*t->tree, pick_stream(t, stream));

int ecfft_enter_dev(const ecfft_tree* t, const void* d_coeffs, size_t n, void* d_evals, void* stream) {
  return guard([&] { DEV_ENGINE eng.enter(dptr(d_coeffs), dptr(d_evals), n); });
}
int ecfft_enter_range_dev(const ecfft_tree* t, const void* d_in, size_t n, size_t m_lo, size_t m_hi, void* d_out, void* stream) {
  return guard([&] { DEV_ENGINE eng.enter_range(dptr(d_in), dptr(d_out), n, m_lo, m_hi); });
}
int ecfft_exit_dev(const ecfft_tree* t, const void* d_evals, size_t n, void* d_coeffs, void* stream) {
  return guard([&] { DEV_ENGINE eng.exit(dptr(d_evals), dptr(d_coeffs), n); });
}
int ecfft_extend_dev(const ecfft_tree* t, const void* d_evals, size_t n, int moiety, void* d_out, void* stream) {
  return guard([&] {
    DEV_ENGINE
    require(moiety == 0 || moiety == 1, ERR_INVALID_ARG, "bad moiety");
    require(n > 0 && n <= ((size_t)1 << 62), ERR_NOT_POW2, "bad length");
    eng.extend(dptr(d_evals), dptr(d_out), n, 1, (Moiety)moiety);
  });
}
int ecfft_mextend_dev(const ecfft_tree* t, const void* d_evals, size_t n, int moiety, void* d_out, void* stream) {
  return guard([&] {
    DEV_ENGINE
    require(moiety == 0 || moiety == 1, ERR_INVALID_ARG, "bad moiety");
    require(n > 0 && n <= ((size_t)1 << 62), ERR_NOT_POW2, "bad length");
    eng.mextend(dptr(d_evals), dptr(d_out), n, (Moiety)moiety, FORM_MONT);
  });
}
int ecfft_degree_dev(const ecfft_tree* t, const void* d_evals, size_t n, size_t* degree, void* stream) {
  return guard([&] {
    DEV_ENGINE
    require(degree != nullptr, ERR_INVALID_ARG, "null degree");
    *degree = eng.degree(dptr(d_evals), n);
  });
}
int ecfft_redc_z0_dev(const ecfft_tree* t, const void* d_evals, const void* d_a, size_t n, void* d_out, void* stream) {
  return guard([&] { DEV_ENGINE eng.redc_user(dptr(d_evals), dptr(d_a), n, S0, dptr(d_out)); });
}
int ecfft_redc_z1_dev(const ecfft_tree* t, const void* d_evals, const void* d_a, size_t n, void* d_out, void* stream) {
  return guard([&] { DEV_ENGINE eng.redc_user(dptr(d_evals), dptr(d_a), n, S1, dptr(d_out)); });
}
int ecfft_modular_reduce_dev(const ecfft_tree* t, const void* d_evals, const void* d_a, const void* d_c, size_t n, void* d_out, void* stream) {
  return guard([&] { DEV_ENGINE eng.mod_user(dptr(d_evals), dptr(d_a), dptr(d_c), n, dptr(d_out)); });
}
int ecfft_pointwise_mul_dev(const ecfft_tree* t, const void* d_a, const void* d_b, size_t n, void* d_out, void* stream) {
  return guard([&] {
    DEV_ENGINE
    if (n) k::mul_mont(dptr(d_out), dptr(d_a), dptr(d_b), n, eng.st);
  });
}
int ecfft_vanish_dev(const ecfft_tree* t, const void* d_domain, size_t n, void* d_out, void* stream) {
  return guard([&] {
    DEV_ENGINE
    require(n > 0 && n <= ((size_t)1 << 62), ERR_NOT_POW2, "bad length");
    eng.vanish(dptr(d_domain), dptr(d_out), n, FORM_MONT);
  });
}

// ---- multi-GPU building blocks (include/ecfft_b200.h) ----------------------------------
static uint32_t log2_exact(size_t v, const char* what) {
  require(v && !(v & (v - 1)), ERR_NOT_POW2, what);
  uint32_t l = 0;
  while (((size_t)1 << l) < v) l++;
  return l;
}
int ecfft_mg_prescale_dev(const ecfft_tree* t, size_t m, size_t pos0, const void* d_in, size_t count, void* d_out, void* stream) {
  return guard([&] {
    DEV_ENGINE
    const Level& lv = eng.level_for(m);
    require(lv.gami[0] != nullptr, ERR_MISSING_TABLES, "normalised tables missing");
    require(pos0 + count <= m / 2, ERR_INVALID_ARG, "slice exceeds the vector");
    k::mul_bcast(dptr(d_out), dptr(d_in), lv.gami[0] + pos0, count, 1, eng.st);
  });
}
int ecfft_mg_cross_dev(const ecfft_tree* t, size_t m, int phase, unsigned j, int role, size_t p_pos0, const void* d_own,
                       const void* d_partner, size_t count, void* d_out, void* stream) {
  return guard([&] {
    DEV_ENGINE
    const Level& lv = eng.level_for(m);
    require((phase == 0 || phase == 1) && (role == 0 || role == 1), ERR_INVALID_ARG, "bad phase/role");
    require(((size_t)2 << j) <= m / 2 && p_pos0 + count <= m / 2, ERR_INVALID_ARG, "level or slice exceeds the vector");
    k::mg_cross(lv, phase, j, role, p_pos0, dptr(d_own), dptr(d_partner), count, dptr(d_out), eng.st);
  });
}
int ecfft_mg_local_dev(const ecfft_tree* t, size_t m, const void* d_in, size_t count, void* d_out, void* stream) {
  return guard([&] {
    DEV_ENGINE
    const Level& lv = eng.level_for(m);
    require(count <= m / 2, ERR_INVALID_ARG, "chunk exceeds the vector");
    k::extend_sub(lv, dptr(d_in), dptr(d_out), log2_exact(count, "chunk length is not a power of two"), eng.st);
  });
}
int ecfft_mg_combine_dev(const ecfft_tree* t, size_t m, size_t i0, const void* d_u0, const void* d_v0, const void* d_u1,
                         const void* d_v1, size_t count, void* d_out, void* stream) {
  return guard([&] {
    DEV_ENGINE
    const Level& lv = eng.level_for(m);
    require(i0 + count <= m / 2, ERR_INVALID_ARG, "slice exceeds the block");
    k::mg_combine(lv, i0, dptr(d_u0), dptr(d_v0), dptr(d_u1), dptr(d_v1), count, dptr(d_out), eng.st);
  });
}


// ---- peer-memory arena (include/ecfft_b200.h "peer exchange") ---------------------------------------
// A flag is one u64 in an arena.  signal: everything enqueued on `stream` before it is visible to every
// GPU of the node once the flag shows `value` (release at system scope).  wait: the stream does not go on
// until the flag (usually in a PEER's arena, read over NVLink) is >= value; a wait that is not satisfied
// within timeout_ms traps, which surfaces as a CUDA error on the next call instead of a hung GPU.
int ecfft_mg_arena_alloc(int device, size_t bytes, void** d_ptr, unsigned char* handle64) {
  return guard([&] {
    require(d_ptr && handle64 && bytes > 0, ERR_INVALID_ARG, "bad arena arguments");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    DeviceGuard dev_guard(device);
    void* p = nullptr;
    ECFFT_CUDA(cudaMalloc(&p, bytes));
    ECFFT_CUDA(cudaMemset(p, 0, bytes));
    cudaIpcMemHandle_t h;
    cudaError_t e = cudaIpcGetMemHandle(&h, p);
    if (e != cudaSuccess) {
      cudaFree(p);
      throw Error(ERR_CUDA, std::string("cudaIpcGetMemHandle: ") + cudaGetErrorString(e));
    }
    memcpy(handle64, &h, 64);
    *d_ptr = p;
  });
}
int ecfft_mg_arena_open(int device, const unsigned char* handle64, void** d_peer_ptr) {
  return guard([&] {
    require(d_peer_ptr && handle64, ERR_INVALID_ARG, "bad arena arguments");
    DeviceGuard dev_guard(device);
    cudaIpcMemHandle_t h;
    memcpy(&h, handle64, 64);
    void* p = nullptr;
    ECFFT_CUDA(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
    *d_peer_ptr = p;
  });
}
int ecfft_mg_arena_close(void* d_peer_ptr) {
  return guard([&] { ECFFT_CUDA(cudaIpcCloseMemHandle(d_peer_ptr)); });
}
int ecfft_mg_arena_free(void* d_ptr) {
  return guard([&] { ECFFT_CUDA(cudaFree(d_ptr)); });
}
int ecfft_mg_arena_status(const void* d_ptr, unsigned long long* status) {
  return guard([&] {
    require(d_ptr != nullptr && status != nullptr, ERR_INVALID_ARG, "null argument");
    ECFFT_CUDA(cudaMemcpy(status, (const unsigned long long*)d_ptr + MG_STATUS_FLAG, sizeof(unsigned long long), cudaMemcpyDeviceToHost));
  });
}
int ecfft_mg_arena_reset(void* d_ptr, void* stream) {
  return guard([&] {
    require(d_ptr != nullptr, ERR_INVALID_ARG, "null arena");
    ECFFT_CUDA(cudaMemsetAsync(d_ptr, 0, MG_FLAG_BYTES, (cudaStream_t)stream));
  });
}
int ecfft_selftest_field(int device, unsigned long long samples, unsigned long long* counters3) {
  return guard([&] {
    require(counters3 != nullptr, ERR_INVALID_ARG, "null output");
    DeviceGuard dev_guard(device);
    unsigned long long* d = nullptr;
    ECFFT_CUDA(cudaMalloc((void**)&d, 3 * sizeof(unsigned long long)));
    cudaError_t e = cudaMemset(d, 0, 3 * sizeof(unsigned long long));
    if (e == cudaSuccess) {
      k::selftest_field(d, samples, nullptr);
      e = cudaMemcpy(counters3, d, 3 * sizeof(unsigned long long), cudaMemcpyDeviceToHost);
    }
    cudaFree(d);
    ECFFT_CUDA(e);
  });
}
int ecfft_mg_arena_bytes(size_t n, int world, size_t* bytes) {
  return guard([&] {
    require(bytes != nullptr, ERR_INVALID_ARG, "null output");
    *bytes = peer_arena_bytes(n, world);
  });
}
int ecfft_enter_peer_dev(const ecfft_tree* t, const void* d_chunk, size_t n, int rank, int world, void* const* arena_bases,
                         unsigned long long epoch, void* d_out_chunk, void* stream) {
  return guard([&] {
    DEV_ENGINE
    require(arena_bases != nullptr, ERR_INVALID_ARG, "null arena table");
    enter_peer(eng, dptr(d_chunk), n, rank, world, arena_bases, epoch, dptr(d_out_chunk));
  });
}
int ecfft_mg_exit_arena_bytes(size_t n, int world, size_t* bytes) {
  return guard([&] {
    require(bytes != nullptr, ERR_INVALID_ARG, "null output");
    *bytes = peer_exit_arena_bytes(n, world);
  });
}
int ecfft_exit_peer_dev(const ecfft_tree* t, const void* d_chunk, size_t n, int rank, int world, void* const* arena_bases,
                        unsigned long long epoch, void* d_out_chunk, void* stream) {
  return guard([&] {
    DEV_ENGINE
    require(arena_bases != nullptr, ERR_INVALID_ARG, "null arena table");
    exit_peer(eng, dptr(d_chunk), n, rank, world, arena_bases, epoch, dptr(d_out_chunk));
  });
}
int ecfft_mg_signal_dev(void* d_flag, unsigned long long value, void* stream) {
  return guard([&] {
    require(d_flag != nullptr && ((uintptr_t)d_flag & 7) == 0, ERR_INVALID_ARG, "flag pointer must be 8-byte aligned");
    k::mg_sync((unsigned long long*)d_flag, value, nullptr, nullptr, 0, (cudaStream_t)stream);
  });
}
int ecfft_mg_wait_dev(const void* d_flag, unsigned long long value, unsigned timeout_ms, void* stream) {
  return guard([&] {
    require(d_flag != nullptr && ((uintptr_t)d_flag & 7) == 0, ERR_INVALID_ARG, "flag pointer must be 8-byte aligned");
    k::mg_sync(nullptr, value, (const unsigned long long*)d_flag, nullptr, timeout_ms, (cudaStream_t)stream);
  });
}

}  // extern "C"
