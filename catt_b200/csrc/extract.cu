// extract.cu — K5: per-record geometric filters, real_score, read clipping and the single-read finder!.
//
// Replaces, for one SAM record (or one V/J record pair of a both-mapped read):
//   assignV (catt.jl:86-106), assignJ (catt.jl:108-132), real_score + HammingDistanceG4 (catt.jl:58-84),
//   the both-mapped clipping of catt.jl:298-305, finder! (Jtool.jl:355-370) and the filters of catt.jl:267-279.
// One warp per item; the SAM-oriented read is unpacked into shared memory once.
#include "finder.cuh"
#include "internal.cuh"

namespace catt {
namespace {

constexpr int EX_WARPS = 4;
constexpr unsigned FULL = 0xffffffffu;

struct ExWarp {
  uint8_t seq[MAX_READ + 8];
  FinderScratch fs;
};

__device__ __forceinline__ uint32_t sam_code(const uint32_t* __restrict__ w, const uint32_t* __restrict__ nm, int len, int i, int strand) {
  const int p = strand ? len - 1 - i : i;
  if ((nm[p >> 5] >> (p & 31)) & 1u) return 4u;
  const uint32_t c = (w[p >> 4] >> ((p & 15) * 2)) & 3u;
  return strand ? 3u - c : c;
}

// real_score (catt.jl:69-84): mismatches of the first M block padded to the read / record ends
__device__ bool real_score_warp(const uint8_t* seq, int rdlen, const uint8_t* __restrict__ ref, int reflen, int start, int term,
                                int r_pos, int allowance, int lane) {
  const int lgt = term - start;
  const int left_pad = min(start, r_pos) - 1;
  const int right_pad = max(0, min(rdlen - term, reflen - 9 - (r_pos + lgt)));
  const int n = lgt + 1 + left_pad + right_pad;
  const int upper = (left_pad + right_pad) / 3 + allowance;
  int cnt = 0;
  for (int base = 0; base < n; base += 32) {
    const int idx = base + lane;
    const bool mm = idx < n && seq[start - left_pad - 1 + idx] != ref[r_pos - left_pad - 1 + idx];
    cnt += __popc(__ballot_sync(FULL, mm));
  }
  return cnt <= upper;
}

__global__ void __launch_bounds__(EX_WARPS * 32) extract_kernel(RefDev vref, RefDev jref, ReadsDev reads, const FinderDev* __restrict__ gfd,
                                                                const ExtractItem* __restrict__ items, ExtractOut* __restrict__ out,
                                                                int64_t n, int kmer, int allow_v, int allow_j) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  FinderDev* fd = reinterpret_cast<FinderDev*>(smem_raw);
  ExWarp* wsall = reinterpret_cast<ExWarp*>(smem_raw + sizeof(FinderDev));
  for (int i = threadIdx.x; i < (int)(sizeof(FinderDev) / 4); i += blockDim.x)
    reinterpret_cast<uint32_t*>(fd)[i] = reinterpret_cast<const uint32_t*>(gfd)[i];
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  ExWarp& ws = wsall[warp];
  for (int64_t it = (int64_t)blockIdx.x * EX_WARPS + warp; it < n; it += (int64_t)gridDim.x * EX_WARPS) {
    const ExtractItem item = items[it];
    const int m = reads.len[item.read];
    const uint32_t* __restrict__ w = reads.words + reads.off[item.read];
    const uint32_t* __restrict__ nm = w + read_base_words(m);
    for (int i = lane; i < m; i += 32) ws.seq[i] = (uint8_t)sam_code(w, nm, m, i, item.strand);
    __syncwarp();
    ExtractOut o;
    o.status = 0; o.seq_off = 0; o.seq_len = 0; o.cdr3_off = -1; o.cdr3_len = 0; o.able = 0; o.pad_ = 0;
    bool keep = true;
    int off = 0, slen = 0;
    const int start = item.qb + 1;
    if (item.kind == 0) {                                            // assignV
      const int r_pos = item.rb + 1, r_lgt = vref.rec_len[item.rid];
      if ((r_lgt - r_pos) - (m - start) > -10) keep = false;
      if (keep && !real_score_warp(ws.seq, m, vref.fa + vref.rec_off[item.rid], r_lgt, start, item.m_end, r_pos, allow_v, lane)) keep = false;
      off = start - 1; slen = m - off;
    } else if (item.kind == 1) {                                     // assignJ
      const int r_pos = item.rb + 1, r_lgt = jref.rec_len[item.rid];
      if (start - 1 < 5) keep = false;
      if (keep && !real_score_warp(ws.seq, m, jref.fa + jref.rec_off[item.rid], r_lgt, start, item.m_end, r_pos, allow_j, lane)) keep = false;
      off = 0; slen = min(m, start + r_lgt);
    } else {                                                         // both-mapped: SAM.sequence(p1) == SAM.sequence(p2), then seq[s:t]
      bool same = true;
      if (item.read2 != item.read || item.j_strand != item.strand) {
        const int m2 = reads.len[item.read2];
        same = (m2 == m);
        if (same) {
          const uint32_t* __restrict__ w2 = reads.words + reads.off[item.read2];
          const uint32_t* __restrict__ nm2 = w2 + read_base_words(m2);
          bool diff = false;
          for (int base = 0; base < m; base += 32) {
            const int i = base + lane;
            diff |= __ballot_sync(FULL, i < m && ws.seq[i] != (uint8_t)sam_code(w2, nm2, m2, i, item.j_strand)) != 0;
          }
          same = !diff;
        }
      }
      off = start - 1; slen = item.j_m_end - start + 1;
      if (!same) { keep = false; o.status = 2; }
      else if (slen < kmer) keep = false;
    }
    if (keep) {
      bool hasn = false;
      for (int base = 0; base < slen; base += 32) {
        const int i = base + lane;
        hasn |= __ballot_sync(FULL, i < slen && ws.seq[off + i] > 3) != 0;
      }
      if (hasn) keep = false;                                        // dropped by catt.jl:279 / 303 whatever finder! says
    }
    if (keep) {
      int lo, l3;
      const bool able = finder_warp(ws.seq + off, slen, *fd, false, ws.fs, lane, &lo, &l3);
      o.able = able;
      if (!able) keep = false;
      if (item.kind != 2 && !(slen > kmer)) keep = false;
      o.cdr3_off = (int16_t)lo; o.cdr3_len = (int16_t)l3;
    }
    o.status = keep ? 1 : (o.status == 2 ? 2 : 0); o.seq_off = (int16_t)off; o.seq_len = (int16_t)slen;
    if (lane == 0) out[it] = o;
    __syncwarp();
  }
}

// finder! on raw reads (parity hook): the whole read is the sequence
__global__ void __launch_bounds__(EX_WARPS * 32) finder_raw_kernel(ReadsDev reads, const FinderDev* __restrict__ gfd, int array_version,
                                                                   int32_t* __restrict__ out3) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  FinderDev* fd = reinterpret_cast<FinderDev*>(smem_raw);
  ExWarp* wsall = reinterpret_cast<ExWarp*>(smem_raw + sizeof(FinderDev));
  for (int i = threadIdx.x; i < (int)(sizeof(FinderDev) / 4); i += blockDim.x)
    reinterpret_cast<uint32_t*>(fd)[i] = reinterpret_cast<const uint32_t*>(gfd)[i];
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  ExWarp& ws = wsall[warp];
  for (int64_t it = (int64_t)blockIdx.x * EX_WARPS + warp; it < reads.n; it += (int64_t)gridDim.x * EX_WARPS) {
    const int m = reads.len[it];
    const uint32_t* __restrict__ w = reads.words + reads.off[it];
    const uint32_t* __restrict__ nm = w + read_base_words(m);
    bool hasn = false;
    for (int base = 0; base < m; base += 32) {
      const int i = base + lane;
      uint32_t c = 0;
      if (i < m) { c = sam_code(w, nm, m, i, 0); ws.seq[i] = (uint8_t)c; }
      hasn |= __ballot_sync(FULL, c > 3) != 0;
    }
    __syncwarp();
    int lo = -1, l3 = 0;
    bool able = false;
    if (!hasn) able = finder_warp(ws.seq, m, *fd, array_version != 0, ws.fs, lane, &lo, &l3);
    if (lane == 0) { out3[3 * it] = able; out3[3 * it + 1] = lo; out3[3 * it + 2] = l3; }
    __syncwarp();
  }
}

// real_score alone (parity hook): one warp per SAM record
__global__ void __launch_bounds__(EX_WARPS * 32) real_score_kernel(RefDev ref, ReadsDev reads, const int32_t* __restrict__ rid,
                                                                   const int32_t* __restrict__ strand, const int32_t* __restrict__ qb,
                                                                   const int32_t* __restrict__ m_end, const int32_t* __restrict__ rb, int allowance,
                                                                   uint8_t* __restrict__ out) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  ExWarp* wsall = reinterpret_cast<ExWarp*>(smem_raw);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  ExWarp& ws = wsall[warp];
  for (int64_t it = (int64_t)blockIdx.x * EX_WARPS + warp; it < reads.n; it += (int64_t)gridDim.x * EX_WARPS) {
    const int m = reads.len[it];
    const uint32_t* __restrict__ w = reads.words + reads.off[it];
    const uint32_t* __restrict__ nm = w + read_base_words(m);
    for (int i = lane; i < m; i += 32) ws.seq[i] = (uint8_t)sam_code(w, nm, m, i, strand[it]);
    __syncwarp();
    const bool ok = real_score_warp(ws.seq, m, ref.fa + ref.rec_off[rid[it]], ref.rec_len[rid[it]], qb[it] + 1, m_end[it], rb[it] + 1, allowance, lane);
    if (lane == 0) out[it] = ok;
    __syncwarp();
  }
}

}  // namespace

int launch_real_score(const RefDev& ref, const ReadsDev& reads, const int32_t* d_rid, const int32_t* d_strand, const int32_t* d_qb, const int32_t* d_m_end,
                      const int32_t* d_rb, int allowance, uint8_t* d_out, int sm_count, cudaStream_t st) {
  if (reads.n == 0) return CATT_OK;
  const size_t smem = sizeof(ExWarp) * EX_WARPS;
  static std::atomic<unsigned long long> optin{0};
  int rc = smem_optin(real_score_kernel, (int)smem, optin);
  if (rc) return rc;
  const int grid = (int)std::min<int64_t>((reads.n + EX_WARPS - 1) / EX_WARPS, (int64_t)sm_count * 8);
  real_score_kernel<<<grid, EX_WARPS * 32, smem, st>>>(ref, reads, d_rid, d_strand, d_qb, d_m_end, d_rb, allowance, d_out);
  CATT_CUDA(cudaGetLastError());
  return CATT_OK;
}

int launch_extract(const RefDev& vref, const RefDev& jref, const ReadsDev& reads, const FinderDev* d_finder,
                   const ExtractItem* d_items, ExtractOut* d_out, int64_t n, int kmer, int allow_v, int allow_j, int sm_count,
                   cudaStream_t st, int* launches) {
  if (n == 0) return CATT_OK;
  const size_t smem = sizeof(FinderDev) + sizeof(ExWarp) * EX_WARPS;
  static std::atomic<unsigned long long> optin{0};
  int rc = smem_optin(extract_kernel, (int)smem, optin);
  if (rc) return rc;
  const int grid = (int)std::min<int64_t>((n + EX_WARPS - 1) / EX_WARPS, (int64_t)sm_count * 8);
  extract_kernel<<<grid, EX_WARPS * 32, smem, st>>>(vref, jref, reads, d_finder, d_items, d_out, n, kmer, allow_v, allow_j);
  CATT_CUDA(cudaGetLastError());
  *launches += 1;
  return CATT_OK;
}

int launch_finder_raw(const ReadsDev& reads, const FinderDev* d_finder, int array_version, int32_t* d_out3, int sm_count, cudaStream_t st) {
  if (reads.n == 0) return CATT_OK;
  const size_t smem = sizeof(FinderDev) + sizeof(ExWarp) * EX_WARPS;
  static std::atomic<unsigned long long> optin{0};
  int rc = smem_optin(finder_raw_kernel, (int)smem, optin);
  if (rc) return rc;
  const int grid = (int)std::min<int64_t>((reads.n + EX_WARPS - 1) / EX_WARPS, (int64_t)sm_count * 8);
  finder_raw_kernel<<<grid, EX_WARPS * 32, smem, st>>>(reads, d_finder, array_version, d_out3);
  CATT_CUDA(cudaGetLastError());
  return CATT_OK;
}

}  // namespace catt
