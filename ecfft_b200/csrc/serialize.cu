// CanonicalSerialize / CanonicalDeserialize of FFTree<Fp> (reference src/fftree.rs:510-660) in the
// ark-serialize 0.4 wire conventions: u64 LE lengths, Fp = 32-byte LE canonical integer (exactly
// the plain-form limbs the device tables hold), fixed arrays unprefixed, bool = 1 byte.
// Field order: f, recombine_matrices, decompose_matrices, rational_maps, xnn_s, z0_s1, z1_s0,
// [uncompressed: xnn_s_inv, z0_inv_s1, z1_inv_s0], z0z0_rem_xnn_s, z1z1_rem_xnn_s, has_subtree,
// [subtree] (fftree.rs:532-552).
#include <string.h>

#include "engine.h"

namespace ecfft {

namespace {
struct Writer {
  uint8_t* p;
  size_t cap, len;
  bool count_only;
  cudaStream_t st;
  void bytes(const void* src, size_t n) {
    if (!count_only) {
      if (len + n > cap) throw Error(ERR_BUFFER_TOO_SMALL, "serialize: buffer too small");
      memcpy(p + len, src, n);
    }
    len += n;
  }
  void u64(uint64_t v) {
    uint8_t b[8];
    for (int i = 0; i < 8; i++) b[i] = (uint8_t)(v >> (8 * i));
    bytes(b, 8);
  }
  void dev(const Fp* d, size_t count) {  // device array -> wire (same bytes)
    if (!count_only && count) {
      if (len + count * 32 > cap) throw Error(ERR_BUFFER_TOO_SMALL, "serialize: buffer too small");
      ECFFT_CUDA(cudaMemcpyAsync(p + len, d, count * 32, cudaMemcpyDeviceToHost, st));
      ECFFT_CUDA(cudaStreamSynchronize(st));
    }
    len += count * 32;
  }
  void vec(const Fp* d, size_t count) {
    u64(count);
    dev(d, count);
  }
  void host_vec(const std::vector<Fp>& v) {
    u64(v.size());
    bytes(v.data(), v.size() * 32);
  }
};

void write_tree(Writer& w, const Tree& t, bool compressed) {
  const size_t n = t.n();
  Engine eng(t, w.st);
  for (int k = (int)t.log_n; k >= 0; k--) {
    const Level& lv = t.levels[k];
    const size_t N = (size_t)1 << k, stride = n / N, h = N / 2;
    if (N > 1 && !lv.has_z) throw Error(ERR_MISSING_TABLES, "serialize: tree was built without the Z tables");
    // f of this chain level: strided view of the top tree's layers (derive_subtree)
    w.u64(2 * N);
    if (w.count_only) {
      w.len += 2 * N * 32;
    } else {
      Fp* fl = eng.tmp(2 * N);
      ECFFT_CUDA(cudaMemsetAsync(fl, 0, sizeof(Fp), w.st));
      for (uint32_t kk = 0; kk <= (uint32_t)k; kk++)
        k::copy_strided(fl + (N >> kk), t.f + (n >> kk), N >> kk, stride, w.st);
      w.dev(fl, 2 * N);
      eng.release(fl);
    }
    w.u64(N);
    w.dev(lv.rmat, 4 * N);
    w.u64(N);
    w.dev(lv.dmat, 4 * N);
    w.u64((uint64_t)k);
    for (int i = 0; i < k; i++) {
      w.host_vec(t.maps[i].num);
      w.host_vec(t.maps[i].den);
    }
    w.vec(lv.xnn_s, N);
    w.vec(lv.z0_s1, h);
    w.vec(lv.z1_s0, h);
    if (!compressed) {
      w.vec(lv.xnn_s_inv, N);
      w.vec(lv.z0_inv_s1, h);
      w.vec(lv.z1_inv_s0, h);
    }
    w.vec(lv.z0z0, N > 1 ? N : 0);
    w.vec(lv.z1z1, N > 1 ? N : 0);
    uint8_t has_sub = k > 0;
    w.bytes(&has_sub, 1);
  }
}

struct Reader {
  const uint8_t* p;
  size_t len, pos;
  uint64_t u64() {
    if (pos + 8 > len) throw Error(ERR_BAD_BYTES, "deserialize: truncated input");
    uint64_t v = 0;
    for (int i = 7; i >= 0; i--) v = (v << 8) | p[pos + i];
    pos += 8;
    return v;
  }
  const uint8_t* take(size_t count) {  // count Fp
    if (count > (len - pos) / 32) throw Error(ERR_BAD_BYTES, "deserialize: truncated input");
    const uint8_t* r = p + pos;
    pos += count * 32;
    return r;
  }
};
}  // namespace

size_t serialized_size(const Tree& t, bool compressed) {
  Writer w{nullptr, 0, 0, true, t.stream};
  write_tree(w, t, compressed);
  return w.len;
}
size_t serialize(const Tree& t, bool compressed, uint8_t* buf, size_t cap) {
  Writer w{buf, cap, 0, false, t.stream};
  write_tree(w, t, compressed);
  return w.len;
}

Tree* deserialize(const uint8_t* buf, size_t len, bool compressed, int device) {
  Reader r{buf, len, 0};
  Tree* t = nullptr;
  unsigned long long* bad = nullptr;
  try {
    int top = -1;
    for (int k = -1;;) {
      uint64_t nf = r.u64();
      if (nf < 2 || (nf & (nf - 1))) throw Error(ERR_BAD_BYTES, "deserialize: f length is not a power of two");
      const size_t N = nf / 2, h = N / 2;
      uint32_t lg = 0;
      while (((size_t)1 << lg) < N) lg++;
      if (!t) {
        if (lg >= 36) throw Error(ERR_BAD_BYTES, "deserialize: tree too large");
        top = (int)lg;
        t = new Tree();
        t->device = device;
        t->log_n = lg;
        t->parts = PARTS_FULL;
        ECFFT_CUDA(cudaSetDevice(device));
        ECFFT_CUDA(cudaStreamCreateWithFlags(&t->stream, cudaStreamNonBlocking));
        t->levels.assign(lg + 1, Level());
        t->maps.assign(lg, RatMapHost());
        ECFFT_CUDA(cudaMalloc((void**)&bad, sizeof *bad));
        t->owned.push_back(bad);
        ECFFT_CUDA(cudaMemsetAsync(bad, 0, sizeof *bad, t->stream));
        k = (int)lg;
      } else if ((int)lg != k) {
        throw Error(ERR_BAD_BYTES, "deserialize: subtree does not have half the leaves");
      }
      cudaStream_t st = t->stream;
      auto upload = [&](size_t count) -> Fp* {
        const uint8_t* src = r.take(count);
        Fp* d = t->dalloc(count);
        if (count) {
          ECFFT_CUDA(cudaMemcpyAsync(d, src, count * 32, cudaMemcpyHostToDevice, st));
          k::count_noncanonical(bad, d, count, st);
        }
        return d;
      };
      auto upload_vec = [&](size_t expect) -> Fp* {
        if (r.u64() != expect) throw Error(ERR_BAD_BYTES, "deserialize: unexpected vector length");
        return upload(expect);
      };
      Level& lv = t->levels[k];
      lv.log_n = (uint32_t)k;
      if (k == top) {
        t->f = upload(2 * N);
      } else {
        r.take(2 * N);  // a subtree's f is the strided view of the top tree's (fftree.rs:465-482)
      }
      if (r.u64() != N) throw Error(ERR_BAD_BYTES, "deserialize: bad recombine_matrices length");
      lv.rmat = upload(4 * N);
      if (r.u64() != N) throw Error(ERR_BAD_BYTES, "deserialize: bad decompose_matrices length");
      lv.dmat = upload(4 * N);
      if (r.u64() != (uint64_t)k) throw Error(ERR_BAD_BYTES, "deserialize: bad rational_maps length");
      for (int i = 0; i < k; i++) {
        RatMapHost m;
        for (int part = 0; part < 2; part++) {
          uint64_t cnt = r.u64();
          if (cnt > 1024) throw Error(ERR_BAD_BYTES, "deserialize: rational map too long");
          const uint8_t* src = r.take(cnt);
          std::vector<Fp>& dst = part == 0 ? m.num : m.den;
          dst.resize(cnt);
          memcpy(dst.data(), src, cnt * 32);
          for (const Fp& c : dst)
            if (!fp_eq(c, fp_canon(c))) throw Error(ERR_BAD_BYTES, "deserialize: element >= p");
        }
        if (k == top) t->maps[i] = m;
      }
      lv.xnn_s = upload_vec(N);
      lv.z0_s1 = upload_vec(h);
      lv.z1_s0 = upload_vec(h);
      if (compressed) {  // fftree.rs:621-628
        lv.xnn_s_inv = t->dalloc(N);
        lv.z0_inv_s1 = t->dalloc(h);
        lv.z1_inv_s0 = t->dalloc(h);
        ECFFT_CUDA(cudaMemcpyAsync(lv.xnn_s_inv, lv.xnn_s, N * 32, cudaMemcpyDeviceToDevice, st));
        k::batch_inverse(lv.xnn_s_inv, N, st);
        if (h) {
          ECFFT_CUDA(cudaMemcpyAsync(lv.z0_inv_s1, lv.z0_s1, h * 32, cudaMemcpyDeviceToDevice, st));
          ECFFT_CUDA(cudaMemcpyAsync(lv.z1_inv_s0, lv.z1_s0, h * 32, cudaMemcpyDeviceToDevice, st));
          k::batch_inverse(lv.z0_inv_s1, h, st);
          k::batch_inverse(lv.z1_inv_s0, h, st);
        }
      } else {
        lv.xnn_s_inv = upload_vec(N);
        lv.z0_inv_s1 = upload_vec(h);
        lv.z1_inv_s0 = upload_vec(h);
      }
      lv.z0z0 = upload_vec(N > 1 ? N : 0);
      lv.z1z1 = upload_vec(N > 1 ? N : 0);
      lv.has_z = N > 1;
      ECFFT_CUDA(cudaStreamSynchronize(st));  // host bytes consumed before moving on
      if (r.pos + 1 > r.len) throw Error(ERR_BAD_BYTES, "deserialize: truncated input");
      uint8_t has_sub = r.p[r.pos++];
      if (has_sub > 1) throw Error(ERR_BAD_BYTES, "deserialize: bad bool");
      if ((has_sub == 1) != (k > 0)) throw Error(ERR_BAD_BYTES, "deserialize: subtree chain does not end at one leaf");
      if (!has_sub) break;
      k--;
    }
    unsigned long long nbad = 0;
    ECFFT_CUDA(cudaMemcpyAsync(&nbad, bad, sizeof nbad, cudaMemcpyDeviceToHost, t->stream));
    ECFFT_CUDA(cudaStreamSynchronize(t->stream));
    if (nbad) throw Error(ERR_BAD_BYTES, "deserialize: element >= p");
    for (uint32_t lvl = 1; lvl <= t->log_n; lvl++) build_norm_tables(*t, lvl);
    ECFFT_CUDA(cudaStreamSynchronize(t->stream));
    const size_t n = t->n();
    t->base_leaf0 = fp_zero();
    t->base_leaf1 = fp_zero();
    if (t->log_n >= 1) {
      ECFFT_CUDA(cudaMemcpyAsync(&t->base_leaf0, t->f + n, sizeof(Fp), cudaMemcpyDeviceToHost, t->stream));
      ECFFT_CUDA(cudaMemcpyAsync(&t->base_leaf1, t->f + n + n / 2, sizeof(Fp), cudaMemcpyDeviceToHost, t->stream));
      ECFFT_CUDA(cudaStreamSynchronize(t->stream));
    }
  } catch (...) {
    delete t;
    throw;
  }
  return t;
}

}  // namespace ecfft
