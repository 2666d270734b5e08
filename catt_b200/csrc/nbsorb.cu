// nbsorb.cu — the two O(n x m) Hamming scans inside the reference's `nbsorb` error correction:
//   * linkRLG for every leaf (Jtool.jl:396-421, driven by the Threads.@threads loop at Jtool.jl:524-531), and
//   * "is any high-confidence sequence within 3 mismatches of this low-count one" (Jtool.jl:545-560).
// Everything else in nbsorb (grouping, the floating-point thresholds, the tree assignment) stays in Julia.
//
// Data layout: every sequence of a group has the same length lgt (nbsorb is called per CDR3 length, catt.jl:624-633) and is over ACGTN.
// A sequence becomes three bit planes of NW = ceil(lgt/32) words each - low code bit, high code bit, "is N" - with the
// positions outside skip+1 .. lgt-skip zeroed in all planes, so the mismatches of HammingDistanceG3(a, b, upper; skip) are
//   popc( (a.lo ^ b.lo) | (a.hi ^ b.hi) | (a.n ^ b.n) )      summed over NW words: 3 LOP3 + 1 POPC per 32 bases.
// One thread owns one query (leaf / low-count sequence) in registers; the targets (roots / high-confidence sequences) stream
// through shared memory in tiles that every thread reads at the same address (broadcast).  gridDim.y splits the targets when
// there are too few queries to fill 148 SMs; partial results meet through atomics.  Integer ALU bound: 4*NW+2 instructions per pair.
#include <cuda_runtime.h>
#include <stdint.h>

#include "host_types.h"
#include "internal.cuh"

namespace catt {
namespace {

constexpr int NB_THREADS = 256;
constexpr int NB_TILE = 128;          // targets per shared-memory tile
constexpr int NB_MAX_LGT = 128;       // the reference calls nbsorb for lgt 21..105 (catt.jl:624)

// thread per (sequence, word): 32 bases -> one word of each plane.  Layout: planes[(seq * 3 + plane) * NW + w].
__global__ void nb_pack_kernel(const char* __restrict__ seqs, int64_t n, int lgt, int nw, int skip, uint32_t* __restrict__ planes,
                               int* __restrict__ bad) {
  const int64_t id = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (id >= n * nw) return;
  const int64_t s = id / nw;
  const int w = (int)(id - s * nw);
  const char* p = seqs + s * lgt;
  uint32_t lo = 0, hi = 0, nn = 0;
  const int i1 = min(lgt, w * 32 + 32);
  for (int i = w * 32; i < i1; i++) {
    const char ch = p[i];
    int code;
    switch (ch) {
      case 'A': code = 0; break;
      case 'C': code = 1; break;
      case 'G': code = 2; break;
      case 'T': code = 3; break;
      case 'N': code = 4; break;
      default: code = 4; atomicExch(bad, 1); break;
    }
    if (i < skip || i >= lgt - skip) continue;           // outside the compared window: identical (zero) in every sequence
    const uint32_t bit = 1u << (i & 31);
    if (code & 1) lo |= bit;
    if (code & 2) hi |= bit;
    if (code & 4) nn |= bit;
  }
  uint32_t* o = planes + s * 3 * nw;
  o[w] = lo; o[nw + w] = hi; o[2 * nw + w] = nn;
}

template <int NW>
__device__ __forceinline__ int mismatches(const uint32_t (&q)[3 * NW], const uint32_t* __restrict__ t) {
  int m = 0;
  #pragma unroll
  for (int w = 0; w < NW; w++) m += __popc((q[w] ^ t[w]) | (q[NW + w] ^ t[NW + w]) | (q[2 * NW + w] ^ t[2 * NW + w]));
  return m;
}

// argmin(@. abs(uplimit[idy] - leaf[2])) (Jtool.jl:408): first minimum, 1-based.  IEEE double subtraction and comparison, as in Julia.
__device__ __forceinline__ int reach_of(const double* __restrict__ u, double cnt) {
  int upper = 1;
  double best = fabs(__dsub_rn(u[0], cnt));
  #pragma unroll
  for (int x = 1; x < 7; x++) {
    const double d = fabs(__dsub_rn(u[x], cnt));
    if (d < best) { best = d; upper = x + 1; }
  }
  return upper;
}

// LINK: count the targets within reach (stop at 2) and remember the first one; !LINK: flag the query if any target is within `upper`.
// PARITY picks the reject test: one parity plane, eight targets per branch (bounds <= 4), or the lo and hi planes, four per branch.
template <int NW, bool LINK, bool PARITY>
__global__ void __launch_bounds__(NB_THREADS)
nb_scan_kernel(const uint32_t* __restrict__ qp, int64_t nq, const uint32_t* __restrict__ tp, int64_t nt, int64_t per_split,
               const double* __restrict__ uplimit, const int64_t* __restrict__ qcount, int upper,
               unsigned int* __restrict__ cnt_out, unsigned int* __restrict__ first_out, uint8_t* __restrict__ flag_out) {
  __shared__ __align__(16) uint32_t tile[NB_TILE * 3 * NW];
  __shared__ __align__(16) uint2 quick2[NB_TILE];         // word 0 of the lo and hi planes, or (PARITY) lo ^ hi packed as uint32
  uint32_t* quick = reinterpret_cast<uint32_t*>(quick2);
  const int64_t q = (int64_t)blockIdx.x * NB_THREADS + threadIdx.x;
  const bool live = q < nq;
  uint32_t my[3 * NW];
  #pragma unroll
  for (int k = 0; k < 3 * NW; k++) my[k] = live ? qp[q * 3 * NW + k] : 0u;
  const double mycount = (LINK && live) ? (double)qcount[q] : 0.0;
  const int bound = LINK ? 7 : upper;                      // reach is 1..7 (Jtool.jl:408), so more than 7 mismatches never link
  const int64_t t_lo = (int64_t)blockIdx.y * per_split, t_hi = min(nt, t_lo + per_split);
  int found = 0;
  unsigned int first = 0xFFFFFFFFu;
  // the full test of target i of the tile; true when this thread needs no further target
  auto full = [&](int i, int64_t t) -> bool {
    const int m = mismatches<NW>(my, tile + i * 3 * NW);
    if (LINK) {
      if (m <= 7 && m <= reach_of(uplimit + t * 7, mycount)) {           // the fp64 path is rare
        if (found == 0) first = (unsigned int)t;
        return ++found >= 2;
      }
      return false;
    }
    if (m <= upper) { found = 1; return true; }
    return false;
  };
  for (int64_t t0 = t_lo; t0 < t_hi; t0 += NB_TILE) {
    const int cntt = (int)min((int64_t)NB_TILE, t_hi - t0);
    // every thread of the block has what it needs (or does not exist): the remaining tiles cannot change any output
    if (__syncthreads_and(!live || found >= (LINK ? 2 : 1))) break;
    for (int k = threadIdx.x; k < cntt * 3 * NW; k += NB_THREADS) tile[k] = tp[t0 * 3 * NW + k];
    for (int k = threadIdx.x; k < NB_TILE; k += NB_THREADS) {      // padding entries: the full test below never looks past cntt
      const uint32_t lo = k < cntt ? tp[(t0 + k) * 3 * NW] : 0u, hi = k < cntt ? tp[(t0 + k) * 3 * NW + NW] : 0u;
      if (PARITY) quick[k] = lo ^ hi; else quick2[k] = make_uint2(lo, hi);
    }
    __syncthreads();
    if (live && found < (LINK ? 2 : 1)) {
      // The reject test is a lower bound of the pair's mismatches from word 0 alone; unrelated sequences fail it and only the rare
      // survivors take the full three-plane test, in target order.  Two planes (popc((lo^lo')|(hi^hi')), 3/4 of unrelated positions
      // count, 4 targets per branch) when the bound is large; for bounds <= 4 the parity plane lo^hi is enough (two bases whose
      // parities differ differ; half of the unrelated positions count): 1 LOP3 + 1 POPC per pair, 8 targets per branch.
      bool done = false;
      if (PARITY) {
        const uint32_t mine = my[0] ^ my[NW];
        for (int i = 0; i < cntt && !done; i += 8) {
          const uint4 a = *reinterpret_cast<const uint4*>(&quick[i]);
          const uint4 b = *reinterpret_cast<const uint4*>(&quick[i + 4]);
          const int m0 = __popc(mine ^ a.x), m1 = __popc(mine ^ a.y), m2 = __popc(mine ^ a.z), m3 = __popc(mine ^ a.w);
          const int m4 = __popc(mine ^ b.x), m5 = __popc(mine ^ b.y), m6 = __popc(mine ^ b.z), m7 = __popc(mine ^ b.w);
          if (min(min(min(m0, m1), min(m2, m3)), min(min(m4, m5), min(m6, m7))) > bound) continue;
          for (int k = 0; k < 8 && i + k < cntt && !done; k++) done = full(i + k, t0 + i + k);
        }
      } else {
        for (int i = 0; i < cntt && !done; i += 4) {
          const uint4 a = *reinterpret_cast<const uint4*>(&quick2[i]);
          const uint4 b = *reinterpret_cast<const uint4*>(&quick2[i + 2]);
          const int m0 = __popc((my[0] ^ a.x) | (my[NW] ^ a.y)), m1 = __popc((my[0] ^ a.z) | (my[NW] ^ a.w));
          const int m2 = __popc((my[0] ^ b.x) | (my[NW] ^ b.y)), m3 = __popc((my[0] ^ b.z) | (my[NW] ^ b.w));
          if (min(min(m0, m1), min(m2, m3)) > bound) continue;
          for (int k = 0; k < 4 && i + k < cntt && !done; k++) done = full(i + k, t0 + i + k);
        }
      }
    }
  }
  if (!live || !found) return;
  if (LINK) { atomicAdd(cnt_out + q, (unsigned int)found); atomicMin(first_out + q, first); }
  else flag_out[q] = 1;
}

// statu = min(matches, 2); holder = 1-based index of the first root within reach, 1 when there is none (Jtool.jl:399, 412-417)
__global__ void nb_link_finish_kernel(const unsigned int* __restrict__ cnt, const unsigned int* __restrict__ first, int64_t n,
                                      int32_t* __restrict__ statu, int32_t* __restrict__ holder) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const unsigned int c = cnt[i];
  statu[i] = (int32_t)min(c, 2u);
  holder[i] = c ? (int32_t)first[i] + 1 : 1;
}

template <bool LINK, bool PARITY>
int launch_scan(int nw, dim3 grid, cudaStream_t st, const uint32_t* qp, int64_t nq, const uint32_t* tp, int64_t nt, int64_t per_split,
                const double* up, const int64_t* qc, int upper, unsigned int* cnt, unsigned int* first, uint8_t* flag) {
  switch (nw) {
    case 1: nb_scan_kernel<1, LINK, PARITY><<<grid, NB_THREADS, 0, st>>>(qp, nq, tp, nt, per_split, up, qc, upper, cnt, first, flag); break;
    case 2: nb_scan_kernel<2, LINK, PARITY><<<grid, NB_THREADS, 0, st>>>(qp, nq, tp, nt, per_split, up, qc, upper, cnt, first, flag); break;
    case 3: nb_scan_kernel<3, LINK, PARITY><<<grid, NB_THREADS, 0, st>>>(qp, nq, tp, nt, per_split, up, qc, upper, cnt, first, flag); break;
    default: nb_scan_kernel<4, LINK, PARITY><<<grid, NB_THREADS, 0, st>>>(qp, nq, tp, nt, per_split, up, qc, upper, cnt, first, flag); break;
  }
  CATT_CUDA(cudaGetLastError());
  return CATT_OK;
}

// Upload + pack both sides.  Buffers: nb[0] raw bytes (queries then targets), nb[1] query planes, nb[2] target planes, nb[3] aux.
int stage_sequences(catt_ctx* c, int lgt, int skip, const char* queries, int64_t nq, const char* targets, int64_t nt, int nw, int* d_bad) {
  int rc;
  if ((rc = c->nb[0].ensure((size_t)(nq + nt) * lgt + 16)) || (rc = c->nb[1].ensure((size_t)nq * 3 * nw * 4 + 16)) ||
      (rc = c->nb[2].ensure((size_t)nt * 3 * nw * 4 + 16))) return rc;
  char* d_raw = c->nb[0].as<char>();
  CATT_CUDA(cudaMemcpyAsync(d_raw, queries, (size_t)nq * lgt, cudaMemcpyHostToDevice, c->stream));
  CATT_CUDA(cudaMemcpyAsync(d_raw + (size_t)nq * lgt, targets, (size_t)nt * lgt, cudaMemcpyHostToDevice, c->stream));
  CATT_CUDA(cudaMemsetAsync(d_bad, 0, sizeof(int), c->stream));
  nb_pack_kernel<<<(unsigned)((nq * nw + 255) / 256), 256, 0, c->stream>>>(d_raw, nq, lgt, nw, skip, c->nb[1].as<uint32_t>(), d_bad);
  nb_pack_kernel<<<(unsigned)((nt * nw + 255) / 256), 256, 0, c->stream>>>(d_raw + (size_t)nq * lgt, nt, lgt, nw, skip, c->nb[2].as<uint32_t>(), d_bad);
  CATT_CUDA(cudaGetLastError());
  return CATT_OK;
}

// how the targets are split over gridDim.y: enough CTAs for two waves of the SMs, at least one tile per split
void split_targets(const catt_ctx* c, int64_t nq, int64_t nt, dim3* grid, int64_t* per_split) {
  const int64_t qblocks = (nq + NB_THREADS - 1) / NB_THREADS;
  int64_t splits = (2 * (int64_t)c->sm_count + qblocks - 1) / qblocks;
  splits = std::max<int64_t>(1, std::min<int64_t>(splits, (nt + NB_TILE - 1) / NB_TILE));
  splits = std::min<int64_t>(splits, 65535);
  *per_split = ((nt + splits - 1) / splits + NB_TILE - 1) / NB_TILE * NB_TILE;
  *grid = dim3((unsigned)qblocks, (unsigned)((nt + *per_split - 1) / *per_split), 1);
}

int check_lgt(int lgt, int skip, const char* who) {
  if (lgt < 1 || lgt > NB_MAX_LGT || skip < 0) {
    set_error("%s: sequence length %d outside 1..%d (nbsorb groups are 21..105 nt, catt.jl:624) or negative skip", who, lgt, NB_MAX_LGT);
    return CATT_E_LIMIT;
  }
  return CATT_OK;
}

}  // namespace
}  // namespace catt

using namespace catt;

extern "C" {

int catt_link_leaves(catt_ctx* c, int32_t lgt, const char* roots, int64_t n_roots, const double* uplimit, const char* leaves,
                     const int64_t* leaf_count, int64_t n_leaves, int32_t* statu, int32_t* root_index) {
  if (!c || n_roots < 0 || n_leaves < 0 || (n_leaves > 0 && (!leaves || !leaf_count || !statu || !root_index)) ||
      (n_roots > 0 && (!roots || !uplimit))) { set_error("catt_link_leaves: bad argument"); return CATT_E_ARG; }
  int rc = check_lgt(lgt, 3, "catt_link_leaves");
  if (rc) return rc;
  if (n_leaves == 0) return CATT_OK;
  if (n_roots == 0) { for (int64_t i = 0; i < n_leaves; i++) { statu[i] = 0; root_index[i] = 1; } return CATT_OK; }
  if (n_roots >= 0x7FFFFFFF) { set_error("catt_link_leaves: too many roots"); return CATT_E_LIMIT; }
  CATT_CUDA(cudaSetDevice(c->device));
  const int nw = (lgt + 31) / 32;
  // aux: [bad flag | uplimit | counts | cnt | first | statu | holder]
  const size_t o_up = 256, o_cnt = o_up + (size_t)n_roots * 7 * 8, o_c = o_cnt + (size_t)n_leaves * 8, o_f = o_c + (size_t)n_leaves * 4,
               o_s = o_f + (size_t)n_leaves * 4, o_h = o_s + (size_t)n_leaves * 4, total = o_h + (size_t)n_leaves * 4;
  if ((rc = c->nb[3].ensure(total))) return rc;
  char* aux = c->nb[3].as<char>();
  int* d_bad = (int*)aux;
  if ((rc = stage_sequences(c, lgt, 3, leaves, n_leaves, roots, n_roots, nw, d_bad))) return rc;
  CATT_CUDA(cudaMemcpyAsync(aux + o_up, uplimit, (size_t)n_roots * 7 * 8, cudaMemcpyHostToDevice, c->stream));
  CATT_CUDA(cudaMemcpyAsync(aux + o_cnt, leaf_count, (size_t)n_leaves * 8, cudaMemcpyHostToDevice, c->stream));
  CATT_CUDA(cudaMemsetAsync(aux + o_c, 0, (size_t)n_leaves * 4, c->stream));
  CATT_CUDA(cudaMemsetAsync(aux + o_f, 0xFF, (size_t)n_leaves * 4, c->stream));
  dim3 grid; int64_t per_split;
  split_targets(c, n_leaves, n_roots, &grid, &per_split);
  if ((rc = launch_scan<true, false>(nw, grid, c->stream, c->nb[1].as<uint32_t>(), n_leaves, c->nb[2].as<uint32_t>(), n_roots, per_split,
                              (const double*)(aux + o_up), (const int64_t*)(aux + o_cnt), 0, (unsigned int*)(aux + o_c),
                              (unsigned int*)(aux + o_f), nullptr))) return rc;
  nb_link_finish_kernel<<<(unsigned)((n_leaves + 255) / 256), 256, 0, c->stream>>>((unsigned int*)(aux + o_c), (unsigned int*)(aux + o_f), n_leaves,
                                                                                   (int32_t*)(aux + o_s), (int32_t*)(aux + o_h));
  CATT_CUDA(cudaGetLastError());
  int bad = 0;
  CATT_CUDA(cudaMemcpyAsync(statu, aux + o_s, (size_t)n_leaves * 4, cudaMemcpyDeviceToHost, c->stream));
  CATT_CUDA(cudaMemcpyAsync(root_index, aux + o_h, (size_t)n_leaves * 4, cudaMemcpyDeviceToHost, c->stream));
  CATT_CUDA(cudaMemcpyAsync(&bad, d_bad, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
  CATT_CUDA(cudaStreamSynchronize(c->stream));
  if (bad) { set_error("catt_link_leaves: sequences must be over ACGTN"); return CATT_E_ARG; }
  return CATT_OK;
}

int catt_near_any(catt_ctx* c, int32_t lgt, const char* queries, int64_t nq, const char* targets, int64_t nt, int32_t upper, int32_t skip,
                  uint8_t* flag) {
  if (!c || nq < 0 || nt < 0 || (nq > 0 && (!queries || !flag)) || (nt > 0 && !targets)) { set_error("catt_near_any: bad argument"); return CATT_E_ARG; }
  int rc = check_lgt(lgt, skip, "catt_near_any");
  if (rc) return rc;
  if (nq == 0) return CATT_OK;
  if (nt == 0) { for (int64_t i = 0; i < nq; i++) flag[i] = 0; return CATT_OK; }
  if (upper < 0) upper = 0;          // HammingDistanceG3 with a negative bound still accepts identical windows (Jtool.jl:69-72)
  CATT_CUDA(cudaSetDevice(c->device));
  const int nw = (lgt + 31) / 32;
  if ((rc = c->nb[3].ensure(256 + (size_t)nq))) return rc;
  char* aux = c->nb[3].as<char>();
  int* d_bad = (int*)aux;
  if ((rc = stage_sequences(c, lgt, skip, queries, nq, targets, nt, nw, d_bad))) return rc;
  CATT_CUDA(cudaMemsetAsync(aux + 256, 0, (size_t)nq, c->stream));
  dim3 grid; int64_t per_split;
  split_targets(c, nq, nt, &grid, &per_split);
  rc = upper <= 4 ? launch_scan<false, true>(nw, grid, c->stream, c->nb[1].as<uint32_t>(), nq, c->nb[2].as<uint32_t>(), nt, per_split, nullptr,
                                             nullptr, upper, nullptr, nullptr, (uint8_t*)(aux + 256))
                  : launch_scan<false, false>(nw, grid, c->stream, c->nb[1].as<uint32_t>(), nq, c->nb[2].as<uint32_t>(), nt, per_split, nullptr,
                                              nullptr, upper, nullptr, nullptr, (uint8_t*)(aux + 256));
  if (rc) return rc;
  int bad = 0;
  CATT_CUDA(cudaMemcpyAsync(flag, aux + 256, (size_t)nq, cudaMemcpyDeviceToHost, c->stream));
  CATT_CUDA(cudaMemcpyAsync(&bad, d_bad, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
  CATT_CUDA(cudaStreamSynchronize(c->stream));
  if (bad) { set_error("catt_near_any: sequences must be over ACGTN"); return CATT_E_ARG; }
  return CATT_OK;
}

}  // extern "C"
