// clonotype.cu — the clonotype table of a result: distinct (CDR3, V, J) with their read counts and merged quality strings.
//
// After catt.jl:581-582 the reference only ever looks at its reads through two groupings:
//   * nbsorb (Jtool.jl:475-490): sort by cdr3, run-length groups, `reduce(MQ, quals)` = per-position maximum of the members' quality
//     characters (Jtool.jl:466-468), count; one call per CDR3 length 21..105 (catt.jl:624-633);
//   * the output table (catt.jl:644): `counter((vs, translate(cdr3), js, cdr3) for it in resRd)`.
// Both are functions of ONE table keyed by (cdr3, V, J) -> (count, merged quality): nbsorb's groups are the sums / maxima over
// the (V, J) of a cdr3, the output rows are the table rows.  It is built here, on the device, from the result lists before they
// leave HBM, so the Julia side can run from a few thousand rows instead of allocating one Myread + strings per returned read
// (SURVEY 8f-2 / 8f-4).  Rows are ordered by (cdr3 length, cdr3, V id, J id), i.e. nbsorb's sort order within each length.
//
// CDR3s carry no N (reads with N never reach the lists, catt.jl:103,124,299) and are 21..105 nt (Jtool.jl:258,304: 6..34 residues
// between C and F), so a CDR3 is a 256-bit key (2 bits per base, left aligned: numeric order == String order over ACGT).
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>

#include "host_types.h"

namespace catt {
namespace {

constexpr int CL_WORDS = 4;                    // 4 x 32 bases
constexpr unsigned FULL = 0xffffffffu;

struct ClonoSrc {                              // where the entries' strings live on the device
  const catt_read* part; const char* d_seq; const char* d_qual; const char* host_seq; const char* host_qual;
};
__device__ __forceinline__ const char* cl_cdr3(const ClonoSrc& S, const catt_read& r) { return S.d_seq + (r.seq - S.host_seq) + r.cdr3_off; }
__device__ __forceinline__ const char* cl_qual(const ClonoSrc& S, const catt_read& r) { return S.d_qual + (r.qual - S.host_qual); }

// per entry: the packed CDR3 (4 words, word-major arrays) and meta = len << 32 | (v + 1) << 16 | (j + 1); len 0 = no usable CDR3
__global__ void clono_key_kernel(ClonoSrc S, int64_t E, unsigned long long* __restrict__ W, unsigned long long* __restrict__ meta,
                                 uint32_t* __restrict__ idx) {
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= E) return;
  const catt_read r = S.part[e];
  unsigned long long w[CL_WORDS] = {0, 0, 0, 0};
  int len = (r.cdr3_off >= 0 && r.cdr3_len > 0 && r.cdr3_len <= 32 * CL_WORDS && r.qual_len >= r.cdr3_len) ? r.cdr3_len : 0;
  if (len) {
    const char* s = cl_cdr3(S, r);
    for (int b = 0; b < len; b++) {
      const uint8_t c = nt_code(s[b]);
      if (c > 3) { len = 0; break; }
      w[b >> 5] |= (unsigned long long)c << (62 - 2 * (b & 31));
    }
  }
  #pragma unroll
  for (int k = 0; k < CL_WORDS; k++) W[(size_t)k * E + e] = len ? w[k] : 0ull;
  meta[e] = ((unsigned long long)len << 32) | ((unsigned long long)(uint16_t)(r.v_id + 1) << 16) | (unsigned long long)(uint16_t)(r.j_id + 1);
  idx[e] = (uint32_t)e;
}
__global__ void clono_gather_kernel(int64_t E, const uint32_t* __restrict__ idx, const unsigned long long* __restrict__ src, int shift,
                                    unsigned long long mask, unsigned long long* __restrict__ out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < E) out[i] = (src[idx[i]] >> shift) & mask;
}
// first entry of every table row (sorted order); entries without a CDR3 (len 0) sort first and form no row
__global__ void clono_flag_kernel(int64_t E, const uint32_t* __restrict__ idx, const unsigned long long* __restrict__ W,
                                  const unsigned long long* __restrict__ meta, uint32_t* __restrict__ flag) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i > E) return;
  if (i == E) { flag[i] = 0; return; }
  const uint32_t a = idx[i];
  bool first = (meta[a] >> 32) != 0;
  if (first && i > 0) {
    const uint32_t b = idx[i - 1];
    bool same = meta[a] == meta[b];
    #pragma unroll
    for (int k = 0; k < CL_WORDS; k++) same = same && W[(size_t)k * E + a] == W[(size_t)k * E + b];
    first = !same;
  }
  flag[i] = first;
}
// row g starts at sorted position start[g]; one word of (len + 1) bytes per row for the two string arenas
__global__ void clono_start_kernel(int64_t E, const uint32_t* __restrict__ flag, const uint32_t* __restrict__ gpos, const uint32_t* __restrict__ idx,
                                   const unsigned long long* __restrict__ meta, uint32_t* __restrict__ start, unsigned long long* __restrict__ rowbytes) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= E || !flag[i]) return;
  start[gpos[i]] = (uint32_t)i;
  rowbytes[gpos[i]] = (meta[idx[i]] >> 32) + 1;
}
// one warp per row: count, the CDR3 text, the per-position maximum of the members' quality characters (reduce(MQ, ...))
__global__ void __launch_bounds__(256) clono_row_kernel(ClonoSrc S, int64_t E, int64_t G, const uint32_t* __restrict__ idx, const uint32_t* __restrict__ start,
                                                        const unsigned long long* __restrict__ roff, catt_clonotype* __restrict__ rows,
                                                        char* __restrict__ seq_out, char* __restrict__ qual_out, const char* host_seq_out,
                                                        const char* host_qual_out) {
  const int lane = threadIdx.x & 31;
  for (int64_t g = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5); g < G; g += (int64_t)gridDim.x * 8) {
    const int64_t s0 = start[g], s1 = g + 1 < G ? (int64_t)start[g + 1] : E;
    const catt_read r0 = S.part[idx[s0]];
    const int len = r0.cdr3_len;
    const char* c3 = cl_cdr3(S, r0);
    char* so = seq_out + roff[g];
    char* qo = qual_out + roff[g];
    for (int p = lane; p <= len; p += 32) {
      char q = 0;
      if (p < len) {
        for (int64_t m = s0; m < s1; m++) { const char x = cl_qual(S, S.part[idx[m]])[p]; q = x > q ? x : q; }
      }
      so[p] = p < len ? c3[p] : 0;
      qo[p] = q;
    }
    if (lane == 0) {
      catt_clonotype row;
      row.cdr3 = host_seq_out + roff[g]; row.qual = host_qual_out + roff[g]; row.cdr3_len = len; row.v_id = r0.v_id; row.j_id = r0.j_id;
      row.reserved_ = 0; row.count = s1 - s0;
      rows[g] = row;
    }
  }
}

template <class T> int need(catt_ctx* c, int slot, int64_t count, T** out) {
  int rc = c->cl[slot].ensure((size_t)std::max<int64_t>(count, 1) * sizeof(T));
  *out = c->cl[slot].as<T>();
  return rc;
}
inline unsigned nblk(int64_t n, int b = 256) { return (unsigned)std::max<int64_t>(1, (n + b - 1) / b); }

}  // namespace

// The E entries of the result (device copies: d_part with HOST string addresses relative to host_seq / host_qual, arenas d_seq /
// d_qual) -> R->clono*.  Runs on the context's stream; the table is tiny next to the lists, so it comes back in three plain copies.
int build_clonotypes(catt_ctx* c, int64_t E, const catt_read* d_part, const char* d_seq, const char* d_qual, const char* host_seq,
                     const char* host_qual, catt_result* R, int* launches) {
  R->n_clono = 0;
  if (E <= 0) return CATT_OK;
  if (E >= 0xFFFFFFF0ll) { set_error("too many reads for the clonotype table"); return CATT_E_LIMIT; }
  cudaStream_t st = c->stream;
  int rc;
  const ClonoSrc S{d_part, d_seq, d_qual, host_seq, host_qual};
  unsigned long long *W, *meta, *ka, *kb, *rowbytes, *roff;
  uint32_t *idx, *idx2, *flag, *gpos, *start;
  if ((rc = need(c, 0, E * CL_WORDS, &W)) || (rc = need(c, 1, E, &meta)) || (rc = need(c, 2, E, &ka)) || (rc = need(c, 3, E, &kb)) ||
      (rc = need(c, 4, E, &idx)) || (rc = need(c, 5, E, &idx2)) || (rc = need(c, 6, E + 1, &flag)) || (rc = need(c, 7, E + 1, &gpos))) return rc;
  clono_key_kernel<<<nblk(E), 256, 0, st>>>(S, E, W, meta, idx);
  CATT_CUDA(cudaGetLastError());
  // LSD passes, least significant key first: (V, J), the four CDR3 words from the last to the first, the length
  struct Pass { const unsigned long long* src; int shift; unsigned long long mask; int bits; };
  const Pass passes[6] = {{meta, 0, 0xFFFFFFFFull, 32}, {W + 3 * E, 0, ~0ull, 64}, {W + 2 * E, 0, ~0ull, 64}, {W + 1 * E, 0, ~0ull, 64},
                          {W, 0, ~0ull, 64}, {meta, 32, 0xFFull, 8}};
  for (const Pass& p : passes) {
    clono_gather_kernel<<<nblk(E), 256, 0, st>>>(E, idx, p.src, p.shift, p.mask, ka);
    CATT_CUDA(cudaGetLastError());
    size_t tmp = 0;
    CATT_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tmp, ka, kb, idx, idx2, E, 0, p.bits, st));
    if ((rc = c->cl[8].ensure(std::max<size_t>(tmp, 256)))) return rc;
    tmp = c->cl[8].cap;
    CATT_CUDA(cub::DeviceRadixSort::SortPairs(c->cl[8].p, tmp, ka, kb, idx, idx2, E, 0, p.bits, st));
    std::swap(idx, idx2);
    *launches += 4;
  }
  clono_flag_kernel<<<nblk(E + 1), 256, 0, st>>>(E, idx, W, meta, flag);
  CATT_CUDA(cudaGetLastError());
  size_t tmp = 0;
  CATT_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, tmp, flag, gpos, E + 1, st));
  if ((rc = c->cl[8].ensure(std::max<size_t>(tmp, 256)))) return rc;
  tmp = c->cl[8].cap;
  CATT_CUDA(cub::DeviceScan::ExclusiveSum(c->cl[8].p, tmp, flag, gpos, E + 1, st));
  uint32_t hG = 0;
  CATT_CUDA(cudaMemcpyAsync(&hG, gpos + E, 4, cudaMemcpyDeviceToHost, st));
  CATT_CUDA(cudaStreamSynchronize(st));
  const int64_t G = hG;
  *launches += 3;
  if (G == 0) return CATT_OK;
  if ((rc = need(c, 9, G + 1, &start)) || (rc = need(c, 10, G + 1, &rowbytes)) || (rc = need(c, 11, G + 1, &roff))) return rc;
  CATT_CUDA(cudaMemsetAsync(rowbytes + G, 0, sizeof(unsigned long long), st));
  clono_start_kernel<<<nblk(E), 256, 0, st>>>(E, flag, gpos, idx, meta, start, rowbytes);
  CATT_CUDA(cudaGetLastError());
  tmp = 0;
  CATT_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, tmp, rowbytes, roff, G + 1, st));
  if ((rc = c->cl[8].ensure(std::max<size_t>(tmp, 256)))) return rc;
  tmp = c->cl[8].cap;
  CATT_CUDA(cub::DeviceScan::ExclusiveSum(c->cl[8].p, tmp, rowbytes, roff, G + 1, st));
  unsigned long long hbytes = 0;
  CATT_CUDA(cudaMemcpyAsync(&hbytes, roff + G, 8, cudaMemcpyDeviceToHost, st));
  CATT_CUDA(cudaStreamSynchronize(st));
  const size_t arena = (size_t)hbytes + 1;
  R->clono.assign((size_t)G, catt_clonotype{});
  R->clono_seq.assign(arena, 0); R->clono_qual.assign(arena, 0);
  catt_clonotype* d_rows; char *d_cs, *d_cq;
  if ((rc = need(c, 12, G, &d_rows)) || (rc = need(c, 13, (int64_t)arena, &d_cs)) || (rc = need(c, 14, (int64_t)arena, &d_cq))) return rc;
  clono_row_kernel<<<(unsigned)std::min<int64_t>((G + 7) / 8, (int64_t)c->sm_count * 16), 256, 0, st>>>(S, E, G, idx, start, roff, d_rows, d_cs, d_cq,
                                                                                                      R->clono_seq.data(), R->clono_qual.data());
  CATT_CUDA(cudaGetLastError());
  CATT_CUDA(cudaMemcpyAsync(R->clono.data(), d_rows, (size_t)G * sizeof(catt_clonotype), cudaMemcpyDeviceToHost, st));
  CATT_CUDA(cudaMemcpyAsync(R->clono_seq.data(), d_cs, arena - 1, cudaMemcpyDeviceToHost, st));
  CATT_CUDA(cudaMemcpyAsync(R->clono_qual.data(), d_cq, arena - 1, cudaMemcpyDeviceToHost, st));
  CATT_CUDA(cudaStreamSynchronize(st));
  R->n_clono = G;
  R->info.bytes_d2h += (int64_t)(G * sizeof(catt_clonotype) + 2 * (arena - 1));
  *launches += 4;
  return CATT_OK;
}

}  // namespace catt
