// On-device corrupted-triple negative sampling.
//   Sampler.sample / RandomModeSampler._sample : skge/sample.py:10-46
//   LCWASampler._sample                        : skge/sample.py:91-110
// The reference keeps the training triples in a Python set and draws with
// numpy.random.randint; here the set is an open-addressing hash table of packed
// 64-bit (s, o, p) keys in HBM and the draws come from Philox4x32-10, so a
// minibatch of negatives is one kernel with no host round trip.  Parity is
// distributional (every emitted negative differs from its positive in exactly
// the sampled slot and is not a training triple), not stream-identical.
#include "common.cuh"

namespace skge {

static constexpr unsigned long long kEmpty = 0xFFFFFFFFFFFFFFFFull;
static constexpr uint32_t kNoEntity = 0xFFFFFFu;  // slot value used for (s, p) pair keys

// 24 bits subject | 24 bits object | 16 bits predicate
__host__ __device__ __forceinline__ unsigned long long pack_key(uint32_t s, uint32_t o, uint32_t p) {
  return ((unsigned long long)p << 48) | ((unsigned long long)s << 24) | (unsigned long long)o;
}
__host__ __device__ __forceinline__ unsigned long long mix64(unsigned long long x) {
  x ^= x >> 30; x *= 0xBF58476D1CE4E5B9ull;
  x ^= x >> 27; x *= 0x94D049BB133111EBull;
  x ^= x >> 31;
  return x;
}

__global__ void tripleset_insert_kernel(unsigned long long *table, unsigned long long mask,
                                        const int32_t *__restrict__ s, const int32_t *__restrict__ o,
                                        const int32_t *__restrict__ p, int64_t T, int pair_only) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < T; i += (int64_t)gridDim.x * blockDim.x) {
    unsigned long long key = pack_key((uint32_t)s[i], pair_only ? kNoEntity : (uint32_t)o[i], (uint32_t)p[i]);
    unsigned long long h = mix64(key) & mask;
    while (true) {
      unsigned long long prev = atomicCAS(table + h, kEmpty, key);
      if (prev == kEmpty || prev == key) break;
      h = (h + 1) & mask;
    }
  }
}

__device__ __forceinline__ bool tripleset_has(const unsigned long long *__restrict__ table,
                                              unsigned long long mask, unsigned long long key) {
  unsigned long long h = mix64(key) & mask;
  while (true) {
    unsigned long long v = __ldg(table + h);
    if (v == key) return true;
    if (v == kEmpty) return false;
    h = (h + 1) & mask;
  }
}

__global__ void tripleset_contains_kernel(const unsigned long long *__restrict__ table, unsigned long long mask,
                                          const int32_t *__restrict__ s, const int32_t *__restrict__ o,
                                          const int32_t *__restrict__ p, int64_t n, uint8_t *__restrict__ out) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    out[i] = tripleset_has(table, mask, pack_key((uint32_t)s[i], (uint32_t)o[i], (uint32_t)p[i]));
}

__device__ __forceinline__ uint4 philox4x32_10(uint4 c, uint2 k) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    uint32_t hi0 = __umulhi(0xD2511F53u, c.x), lo0 = 0xD2511F53u * c.x;
    uint32_t hi1 = __umulhi(0xCD9E8D57u, c.z), lo1 = 0xCD9E8D57u * c.z;
    c = make_uint4(hi1 ^ c.y ^ k.x, lo1, hi0 ^ c.w ^ k.y, lo0);
    k.x += 0x9E3779B9u;
    k.y += 0xBB67AE85u;
  }
  return c;
}

struct SampleArgs {
  const unsigned long long *table, *sp_table;
  unsigned long long mask, sp_mask;
  const int32_t *s, *o, *p, *batch_idx;
  int64_t B;
  int n_per, nmodes, modes[3];
  uint32_t sz[3];
  int ntries;
  uint64_t seed, offset;
  const uint64_t *offset_dev;
  int32_t *out_sp, *out_op, *out_pp, *out_sn, *out_on, *out_pn;
  uint8_t *out_valid;
};

// One thread per emitted pair.
__global__ void __launch_bounds__(256) sample_corrupt_kernel(SampleArgs a) {
  int64_t total = a.B * a.n_per * a.nmodes;
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
    int64_t b = t / (a.n_per * a.nmodes);
    int mode = a.modes[t % a.nmodes];
    int64_t src = a.batch_idx ? a.batch_idx[b] : b;
    uint32_t x[3] = {(uint32_t)a.s[src], (uint32_t)a.o[src], (uint32_t)a.p[src]};
    uint32_t nx[3] = {x[0], x[1], x[2]};
    bool ok = false;
    uint64_t ctr = a.offset + (a.offset_dev ? *a.offset_dev : 0ull) + (uint64_t)t;
    uint2 key = make_uint2((uint32_t)a.seed, (uint32_t)(a.seed >> 32));
    for (int tr = 0; tr < a.ntries && !ok; tr += 4) {
      uint4 r = philox4x32_10(make_uint4((uint32_t)ctr, (uint32_t)(ctr >> 32), (uint32_t)(tr >> 2), 0x5a3c9e1du), key);
      uint32_t rv[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        if (ok || tr + q >= a.ntries) break;
        nx[mode] = (uint32_t)(((unsigned long long)rv[q] * a.sz[mode]) >> 32);  // randint(sz[mode]): sample.py:42
        bool reject = tripleset_has(a.table, a.mask, pack_key(nx[0], nx[1], nx[2]));
        if (!reject && a.sp_table)  // LCWA: (s', p') must have been seen: sample.py:106
          reject = !tripleset_has(a.sp_table, a.sp_mask, pack_key(nx[0], kNoEntity, nx[2]));
        ok = !reject;
      }
    }
    a.out_sp[t] = (int32_t)x[0]; a.out_op[t] = (int32_t)x[1]; a.out_pp[t] = (int32_t)x[2];
    a.out_sn[t] = (int32_t)nx[0]; a.out_on[t] = (int32_t)nx[1]; a.out_pn[t] = (int32_t)nx[2];
    a.out_valid[t] = ok;
  }
}

static bool table_mask(size_t bytes, unsigned long long *mask) {
  size_t cap = bytes / 8;
  if (cap < 16 || (cap & (cap - 1))) return false;
  *mask = cap - 1;
  return true;
}

}  // namespace skge

using namespace skge;

extern "C" {

size_t skge_tripleset_bytes(int64_t T) {
  size_t cap = 1024;
  while (cap < (size_t)(2 * (T > 0 ? T : 1))) cap <<= 1;
  return cap * 8;
}

int skge_tripleset_build(void *table, size_t table_bytes, const int32_t *s, const int32_t *o,
                         const int32_t *p, int64_t T, int pair_keys_only, skge_stream_t stream) {
  unsigned long long mask;
  SKGE_REQUIRE(table && s && o && p && T >= 0, "null argument");
  SKGE_REQUIRE(table_mask(table_bytes, &mask), "table_bytes must be 8 * a power of two");
  SKGE_REQUIRE((size_t)T * 2 <= table_bytes / 8 || T < 512, "table too small (need load factor <= 0.5)");
  cudaStream_t st = as_stream(stream);
  SKGE_CUDA(cudaMemsetAsync(table, 0xFF, table_bytes, st));
  if (T == 0) return 0;
  int64_t blocks = (T + 255) / 256;
  if (blocks > kNumSMs * 16) blocks = kNumSMs * 16;
  tripleset_insert_kernel<<<(int)blocks, 256, 0, st>>>(static_cast<unsigned long long *>(table), mask, s, o, p,
                                                      T, pair_keys_only);
  SKGE_LAUNCH_CHECK();
  return 0;
}

int skge_tripleset_contains(const void *table, size_t table_bytes, const int32_t *s, const int32_t *o,
                            const int32_t *p, int64_t n, uint8_t *out, skge_stream_t stream) {
  unsigned long long mask;
  SKGE_REQUIRE(table && s && o && p && out && n >= 0, "null argument");
  SKGE_REQUIRE(table_mask(table_bytes, &mask), "table_bytes must be 8 * a power of two");
  if (n == 0) return 0;
  int64_t blocks = (n + 255) / 256;
  if (blocks > kNumSMs * 16) blocks = kNumSMs * 16;
  tripleset_contains_kernel<<<(int)blocks, 256, 0, as_stream(stream)>>>(
      static_cast<const unsigned long long *>(table), mask, s, o, p, n, out);
  SKGE_LAUNCH_CHECK();
  return 0;
}

int skge_sample_corrupt(const void *table, size_t table_bytes, const void *sp_table,
                        size_t sp_table_bytes, const int32_t *s, const int32_t *o,
                        const int32_t *p, const int32_t *batch_idx, int64_t B, int n_per,
                        int modes_mask, int64_t N, int64_t M, int ntries, uint64_t seed,
                        uint64_t offset, const uint64_t *offset_dev, int32_t *out_sp, int32_t *out_op,
                        int32_t *out_pp, int32_t *out_sn, int32_t *out_on, int32_t *out_pn,
                        uint8_t *out_valid, skge_stream_t stream) {
  SampleArgs a;
  SKGE_REQUIRE(table && s && o && p && out_sp && out_op && out_pp && out_sn && out_on && out_pn && out_valid,
               "null argument");
  SKGE_REQUIRE(table_mask(table_bytes, &a.mask), "table_bytes must be 8 * a power of two");
  SKGE_REQUIRE(N > 0 && N < (1 << 24) - 1 && M > 0 && M <= (1 << 16), "key packing needs N < 2^24-1, M <= 2^16");
  SKGE_REQUIRE(B >= 0 && n_per > 0 && ntries > 0 && (modes_mask & 7) != 0 && !(modes_mask & ~7), "bad sampling config");
  a.sp_table = nullptr;
  a.sp_mask = 0;
  if (sp_table) {
    SKGE_REQUIRE(table_mask(sp_table_bytes, &a.sp_mask), "sp_table_bytes must be 8 * a power of two");
    a.sp_table = static_cast<const unsigned long long *>(sp_table);
  }
  if (B == 0) return 0;
  a.table = static_cast<const unsigned long long *>(table);
  a.s = s; a.o = o; a.p = p; a.batch_idx = batch_idx;
  a.B = B; a.n_per = n_per;
  a.nmodes = 0;
  for (int m = 0; m < 3; ++m)
    if (modes_mask & (1 << m)) a.modes[a.nmodes++] = m;
  a.sz[0] = (uint32_t)N; a.sz[1] = (uint32_t)N; a.sz[2] = (uint32_t)M;  // sz = (N, N, M): skge/base.py:496
  a.ntries = ntries; a.seed = seed; a.offset = offset; a.offset_dev = offset_dev;
  a.out_sp = out_sp; a.out_op = out_op; a.out_pp = out_pp;
  a.out_sn = out_sn; a.out_on = out_on; a.out_pn = out_pn; a.out_valid = out_valid;
  int64_t total = B * n_per * a.nmodes;
  int64_t blocks = (total + 255) / 256;
  if (blocks > kNumSMs * 16) blocks = kNumSMs * 16;
  sample_corrupt_kernel<<<(int)blocks, 256, 0, as_stream(stream)>>>(a);
  SKGE_LAUNCH_CHECK();
  return 0;
}

}  // extern "C"
