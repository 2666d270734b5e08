// seed.cu — K1: k-mer seeding of reads against the 2-strand V or J text.
//
// Replaces the seeding half of `bwa mem -k 10` inside map2align (catt.jl:37-56): for every read, the set of
// (record, strand) pairs that bwa-mem would extend.  Restated per SURVEY.md Appendix A3 from bwa-0.7.17's
// mem_collect_intv (SMEMs, re-seeding, LAST-like tiling) and mem_chain's bridging discard:
//   RUNS   = maximal exact-match runs of length >= 10 between the read and T = cat + revcomp(cat);
//   pass 1 = runs whose query interval is not strictly inside another run's (the SMEM occurrences);
//   pass 2 = for SMEMs of length >= 15 with <= 10 occurrences: maximal intervals around the middle that occur
//            at least occ+1 times (length >= 10), all occurrences;
//   pass 3 = phase-0 tiling: shortest prefix of length >= 11 with < 20 occurrences, all occurrences;
//   a seed whose reference span crosses a record or the strand boundary is dropped.
// One warp per read; run list, per-position arrays and the candidate bitmap live in shared memory (a global
// scratch variant re-runs the rare reads whose run list overflows it).  The packed reads are read with
// coalesced word loads; the 2-bit text is staged into shared memory by TMA when it fits; the 10-mer bitmap
// (128 KB) and CSR index are probed through L1/L2.
#include "internal.cuh"
#include "tma.cuh"

namespace catt {
namespace {

constexpr int SEED_WARPS = 8;
constexpr int SEED_GRAB = 4;
constexpr int RUNS_SMEM = 256;
constexpr int RUNS_GLOBAL = 16384;
constexpr unsigned FULL = 0xffffffffu;
constexpr int T2_SMEM_MAX_BYTES = 56 * 1024;

// per-warp scratch, carved from dynamic shared memory; the per-position arrays are sized by the longest read of the batch
struct WarpScratch {
  uint32_t* M1;         // [ml] max run end per start position
  uint32_t* Pm;         // [ml] exclusive prefix max of M1; reused as the interval list of pass 2
  uint32_t* cnt;        // [ml] occurrences of the SMEM that starts here
  uint16_t* starts;     // [ml] distinct starts of the runs covering `mid` (pass 2)
  uint8_t* codes;       // [ml]
  uint32_t* startbits;  // [MAX_READ / 32]
  uint32_t* cand;       // [MAX_CW]
  uint32_t* canon;      // [MAX_CW] cand with byte-identical records folded onto their first copy
  int* ctr;             // nruns, ncov, niv
  uint32_t* runs_q;     // [RUNS_SMEM] qs | qe << 16
  uint32_t* runs_t;     // [RUNS_SMEM] start in T
  uint16_t* cov;        // [RUNS_SMEM]
  uint16_t* ord;        // [RUNS_SMEM] run indices by descending query end
};
__host__ __device__ inline size_t warp_scratch_bytes(int ml) {
  return (size_t)ml * (4 * 3 + 2 + 1) + (MAX_READ / 32) * 4 + MAX_CW * 8 + 16 + (size_t)RUNS_SMEM * 12;
}
__device__ inline WarpScratch carve(unsigned char* base, int ml) {
  WarpScratch ws;
  ws.M1 = reinterpret_cast<uint32_t*>(base); base += 4 * ml;
  ws.Pm = reinterpret_cast<uint32_t*>(base); base += 4 * ml;
  ws.cnt = reinterpret_cast<uint32_t*>(base); base += 4 * ml;
  ws.startbits = reinterpret_cast<uint32_t*>(base); base += (MAX_READ / 32) * 4;
  ws.cand = reinterpret_cast<uint32_t*>(base); base += MAX_CW * 4;
  ws.canon = reinterpret_cast<uint32_t*>(base); base += MAX_CW * 4;
  ws.ctr = reinterpret_cast<int*>(base); base += 16;
  ws.runs_q = reinterpret_cast<uint32_t*>(base); base += RUNS_SMEM * 4;
  ws.runs_t = reinterpret_cast<uint32_t*>(base); base += RUNS_SMEM * 4;
  ws.cov = reinterpret_cast<uint16_t*>(base); base += RUNS_SMEM * 2;
  ws.ord = reinterpret_cast<uint16_t*>(base); base += RUNS_SMEM * 2;
  ws.starts = reinterpret_cast<uint16_t*>(base); base += 2 * ml;
  ws.codes = base;
  return ws;
}

struct Scratch {
  uint32_t* runs_q; uint32_t* runs_t; uint16_t* cov; uint16_t* ord; int cap;
};

__device__ __forceinline__ uint32_t tbase(const uint32_t* T2, int p) { return (T2[p >> 4] >> ((p & 15) * 2)) & 3u; }

__device__ __forceinline__ void mark_seed(const RefDev& ref, uint32_t* cand, int rb, int len) {
  const int L = ref.L, re = rb + len;
  if (rb < L && re > L) return;                       // bridges the forward/reverse boundary
  const int b = rb < L ? rb : 2 * L - 1 - rb;
  const int e = (re - 1) < L ? (re - 1) : 2 * L - 1 - (re - 1);
  const int r1 = ref.pos2rid[b], r2 = ref.pos2rid[e];
  if (r1 != r2) return;                               // bridges two records
  const int bit = r1 * 2 + (rb >= L);
  atomicOr(&cand[bit >> 5], 1u << (bit & 31));
}

__device__ __forceinline__ uint32_t warp_sum(uint32_t v) {
  #pragma unroll
  for (int d = 16; d; d >>= 1) v += __shfl_xor_sync(FULL, v, d);
  return v;
}

// pass 2 for one SMEM: maximal intervals containing `mid` with >= m covering runs.  The runs are visited by descending query end
// (sc.ord), so the covering list comes out in that order and "the m-th largest end among the covering runs that start at or
// before sp" is a short counting scan per start instead of a top-m selection over the whole list.
__device__ void reseed(const RefDev& ref, const WarpScratch& ws, const Scratch& sc, int n, int mid, int m, int lane) {
  if (lane == 0) ws.ctr[2] = 0;
  if (lane < MAX_READ / 32) ws.startbits[lane] = 0;
  __syncwarp();
  int nc = 0;
  #pragma unroll 1
  for (int j = 0; j < n; j += 32) {
    const int k = j + lane < n ? (int)sc.ord[j + lane] : -1;
    const uint32_t q = k >= 0 ? sc.runs_q[k] : 0u;
    const int qs = q & 0xffff, qe = q >> 16;
    if (__shfl_sync(FULL, qe, 0) <= mid) break;           // every later run ends at or before mid as well
    const bool cv = k >= 0 && qs <= mid && qe > mid;
    const unsigned bal = __ballot_sync(FULL, cv);
    if (cv) {
      sc.cov[nc + __popc(bal & ((1u << lane) - 1))] = (uint16_t)k;      // cov capacity == run capacity
      atomicOr(&ws.startbits[qs >> 5], 1u << (qs & 31));
    }
    nc += __popc(bal);
  }
  __syncwarp();
  if (nc < m) return;
  int ns = 0;
  #pragma unroll 1
  for (int w = 0; w < MAX_READ / 32; w++) {
    const uint32_t bits = ws.startbits[w];
    const int c = __popc(bits);
    if (lane < c) ws.starts[ns + lane] = (uint16_t)(w * 32 + __fns(bits, 0, lane + 1));
    ns += c;
  }
  __syncwarp();
  uint32_t prev_run = 0;
  #pragma unroll 1
  for (int sb = 0; sb < ns; sb += 32) {
    const int si = sb + lane;
    const bool valid = si < ns;
    const int sp = valid ? ws.starts[si] : -1;
    // e = m-th largest end among the covering runs with qs <= sp (0 when there are fewer than m)
    uint32_t e = 0;
    int c = 0;
    #pragma unroll 1
    for (int k = 0; k < nc; k++) {
      const uint32_t q = sc.runs_q[sc.cov[k]];
      if ((int)(q & 0xffff) <= sp && ++c == m) e = q >> 16;
      if (__all_sync(FULL, c >= m || !valid)) break;
    }
    uint32_t incl = e;
    #pragma unroll
    for (int d = 1; d < 32; d <<= 1) { const uint32_t t = __shfl_up_sync(FULL, incl, d); if (lane >= d) incl = max(incl, t); }
    uint32_t excl = __shfl_up_sync(FULL, incl, 1);
    if (lane == 0) excl = 0;
    excl = max(excl, prev_run);
    if (valid && e > excl && (int)e - sp >= SEED_K) {
      const int idx = atomicAdd(&ws.ctr[2], 1);
      ws.Pm[idx] = (uint32_t)sp | (e << 16);
    }
    prev_run = max(prev_run, __shfl_sync(FULL, incl, 31));
  }
  __syncwarp();
  const int nv = ws.ctr[2];
  #pragma unroll 1
  for (int v = 0; v < nv; v++) {
    const int s = ws.Pm[v] & 0xffff, e = ws.Pm[v] >> 16;
    #pragma unroll 1
    for (int k = lane; k < nc; k += 32) {
      const int r = sc.cov[k];
      const uint32_t q = sc.runs_q[r];
      const int qs = q & 0xffff, qe = q >> 16;
      if (qs <= s && qe >= e) mark_seed(ref, ws.cand, (int)sc.runs_t[r] + (s - qs), e - s);
    }
  }
  __syncwarp();
}

// returns false when the run list overflowed the scratch capacity
__device__ bool seed_one_read(const RefDev& ref, const uint32_t* __restrict__ T2, const ReadsDev& reads, int64_t r,
                              const WarpScratch& ws, const Scratch& sc, int lane) {
  const int len = reads.len[r];
  const uint32_t* __restrict__ w = reads.words + reads.off[r];
  const uint32_t* __restrict__ nm = w + read_base_words(len);
  const int n2 = 2 * ref.L;
  #pragma unroll 1
  for (int i = lane; i < len; i += 32) {
    const uint32_t c = (w[i >> 4] >> ((i & 15) * 2)) & 3u;
    ws.codes[i] = ((nm[i >> 5] >> (i & 31)) & 1u) ? 4 : (uint8_t)c;
    ws.M1[i] = 0; ws.cnt[i] = 0;
  }
  if (lane < MAX_CW) ws.cand[lane] = 0;
  if (lane == 0) ws.ctr[0] = 0;
  __syncwarp();
  // ---- maximal runs >= 10 (each found once, from its left-maximal 10-mer window)
  // step 1: 32 windows at a time.  Every lane looks its window up (bitmap, CSR range, the group to skip); the occurrences of all
  // 32 windows are then expanded by ALL lanes (entry t of the concatenated lists belongs to the lane found by a 5-step search over
  // the warp's prefix sums), so a window with dozens of occurrences - a 10-mer shared by many V alleles - does not serialise
  // on one lane while the others idle, and the position loads of a step are all in flight together.
  int nq_total = 0;
  #pragma unroll 1
  for (int s0 = 0; s0 + SEED_K <= len; s0 += 32) {
    const int s = s0 + lane;
    uint32_t h0 = 0, ex0 = 0, ex1 = 0, cnt = 0;
    if (s + SEED_K <= len) {
      const uint64_t bw = ((uint64_t)w[(s >> 4) + 1] << 32) | w[s >> 4];
      const uint32_t key = (uint32_t)(bw >> ((s & 15) * 2)) & 0xFFFFFu;
      const uint64_t mw = ((uint64_t)nm[(s >> 5) + 1] << 32) | nm[s >> 5];
      if (!((uint32_t)(mw >> (s & 31)) & 0x3FFu) &&                      // no N in the window
          ((__ldg(&ref.bitmap[key >> 5]) >> (key & 31)) & 1u)) {
        h0 = __ldg(&ref.kstart[key]);
        const uint32_t h1 = __ldg(&ref.kstart[key + 1]);
        // occurrences whose preceding text base equals the preceding read base continue a run found earlier: that whole
        // group of the (grouped) occurrence list is skipped
        const uint32_t prev = s > 0 ? ws.codes[s - 1] : 4u;
        ex0 = h1; ex1 = h1;
        if (prev < 4) {
          const unsigned long long sub = __ldg(&ref.ksub[key]);
          ex0 = h0 + (prev ? (uint32_t)(sub >> (16 * (prev - 1))) & 0xFFFFu : 0u);
          ex1 = h0 + ((uint32_t)(sub >> (16 * prev)) & 0xFFFFu);
        }
        cnt = (h1 - h0) - (ex1 - ex0);
      }
    }
    uint32_t incl = cnt;
    #pragma unroll
    for (int d = 1; d < 32; d <<= 1) { const uint32_t t = __shfl_up_sync(FULL, incl, d); if (lane >= d) incl += t; }
    const uint32_t excl = incl - cnt, total = __shfl_sync(FULL, incl, 31);
    #pragma unroll 1
    for (uint32_t t0 = 0; t0 < total; t0 += 32) {
      const uint32_t t = t0 + lane;
      int owner = 0;                                         // the last lane whose exclusive prefix is <= t
      #pragma unroll
      for (int d = 16; d; d >>= 1) { const uint32_t v = __shfl_sync(FULL, excl, owner + d); if (v <= t) owner += d; }
      const uint32_t oh0 = __shfl_sync(FULL, h0, owner), oex0 = __shfl_sync(FULL, ex0, owner), oex1 = __shfl_sync(FULL, ex1, owner);
      const uint32_t oexcl = __shfl_sync(FULL, excl, owner);
      if (t < total) {
        uint32_t h = oh0 + (t - oexcl);
        if (h >= oex0) h += oex1 - oex0;
        const int idx = nq_total + (int)t;
        if (idx < sc.cap) { sc.runs_q[idx] = (uint32_t)(s0 + owner); sc.runs_t[idx] = __ldg(&ref.kpos[h]); }
      }
    }
    nq_total += (int)total;
  }
  if (lane == 0) ws.ctr[0] = nq_total;
  __syncwarp();
  // step 2: the queued runs are extended by all lanes, 16 bases per step on the packed words
  {
    const int nq = min(ws.ctr[0], sc.cap);
    #pragma unroll 1
    for (int k = lane; k < nq; k += 32) {
      const int s = (int)sc.runs_q[k], p = (int)sc.runs_t[k];
      int e = SEED_K;
      const int lim = min(len - s, n2 - p);
      #pragma unroll 1
      while (e < lim) {
        const int a = s + e, b = p + e;
        const uint32_t ra = (uint32_t)((((uint64_t)w[(a >> 4) + 1] << 32) | w[a >> 4]) >> ((a & 15) * 2));
        const uint32_t tb = (uint32_t)((((uint64_t)T2[(b >> 4) + 1] << 32) | T2[b >> 4]) >> ((b & 15) * 2));
        const uint32_t na = (uint32_t)((((uint64_t)nm[(a >> 5) + 1] << 32) | nm[a >> 5]) >> (a & 31)) & 0xFFFFu;
        const uint32_t x = ra ^ tb;
        int adv = x ? (__ffs(x) - 1) >> 1 : 16;
        if (na) adv = min(adv, __ffs(na) - 1);
        e += adv;
        if (adv < 16) break;
      }
      e = min(e, lim);
      sc.runs_q[k] = (uint32_t)s | ((uint32_t)(s + e) << 16);
    }
  }
  __syncwarp();
  const int n = ws.ctr[0];
  if (n > sc.cap) return false;
  if (n == 0) return true;
  // ---- pass 1: SMEM occurrences
  #pragma unroll 1
  for (int k = lane; k < n; k += 32) { const uint32_t q = sc.runs_q[k]; atomicMax(&ws.M1[q & 0xffff], q >> 16); }
  __syncwarp();
  {
    uint32_t carry = 0;
    #pragma unroll 1
    for (int base = 0; base < len; base += 32) {
      const int i = base + lane;
      uint32_t incl = i < len ? ws.M1[i] : 0;
      #pragma unroll
      for (int d = 1; d < 32; d <<= 1) { const uint32_t t = __shfl_up_sync(FULL, incl, d); if (lane >= d) incl = max(incl, t); }
      uint32_t excl = __shfl_up_sync(FULL, incl, 1);
      if (lane == 0) excl = 0;
      excl = max(excl, carry);
      if (i < len) ws.Pm[i] = excl;
      carry = max(carry, __shfl_sync(FULL, incl, 31));
    }
  }
  __syncwarp();
  #pragma unroll 1
  for (int k = lane; k < n; k += 32) {
    const uint32_t q = sc.runs_q[k];
    const uint32_t qs = q & 0xffff, qe = q >> 16;
    if (qe == ws.M1[qs] && ws.Pm[qs] < qe) { mark_seed(ref, ws.cand, (int)sc.runs_t[k], (int)(qe - qs)); atomicAdd(&ws.cnt[qs], 1u); }
  }
  __syncwarp();
  // ---- run indices by descending query end (counting sort over the ends; Pm is free between pass 1 and pass 2's interval list)
  #pragma unroll 1
  for (int i = lane; i < len; i += 32) ws.Pm[i] = 0;
  __syncwarp();
  #pragma unroll 1
  for (int k = lane; k < n; k += 32) atomicAdd(&ws.Pm[len - (int)(sc.runs_q[k] >> 16)], 1u);      // bin 0 = runs that end at the read's end
  __syncwarp();
  {
    uint32_t carry = 0;
    #pragma unroll 1
    for (int base = 0; base < len; base += 32) {
      const int i = base + lane;
      const uint32_t v = i < len ? ws.Pm[i] : 0;
      uint32_t incl = v;
      #pragma unroll
      for (int d = 1; d < 32; d <<= 1) { const uint32_t t = __shfl_up_sync(FULL, incl, d); if (lane >= d) incl += t; }
      if (i < len) ws.Pm[i] = carry + incl - v;
      carry += __shfl_sync(FULL, incl, 31);
    }
  }
  __syncwarp();
  #pragma unroll 1
  for (int k = lane; k < n; k += 32) sc.ord[atomicAdd(&ws.Pm[len - (int)(sc.runs_q[k] >> 16)], 1u)] = (uint16_t)k;
  __syncwarp();
  // ---- pass 2: re-seeding inside long, rare SMEMs
  #pragma unroll 1
  for (int base = 0; base < len; base += 32) {
    const int i = base + lane;
    const bool sel = i < len && ws.cnt[i] > 0 && ws.cnt[i] <= 10 && (int)ws.M1[i] - i >= 15;
    uint32_t bal = __ballot_sync(FULL, sel);
    while (bal) {
      const int bit = __ffs(bal) - 1;
      bal &= bal - 1;
      const int a = base + bit, b = (int)ws.M1[a], m = (int)ws.cnt[a] + 1;
      reseed(ref, ws, sc, n, (a + b) >> 1, m, lane);
    }
  }
  // ---- pass 3: tiles.  Tile [x, e): the shortest e >= x + 11 with fewer than 20 runs covering it.  The number of covering runs
  // is non-increasing in e, so e - 1 is the 20th largest end among the runs that start at or before x (when there are 20 with
  // end >= x + 11): one counting scan over the runs by descending end, no bisection.
  // c11[x] = runs covering [x, x + 11), for every x at once (difference array + prefix sum, in Pm): most tiles of a read lie in
  // its non-V/J part and are covered by nothing - they advance without looking at the run list at all.
  #pragma unroll 1
  for (int i = lane; i < len; i += 32) ws.Pm[i] = 0;
  __syncwarp();
  #pragma unroll 1
  for (int k = lane; k < n; k += 32) {
    const uint32_t q = sc.runs_q[k];
    const int qs = q & 0xffff, qe = q >> 16;
    if (qe - qs >= 11) { atomicAdd(&ws.Pm[qs], 1u); if (qe - 10 < len) atomicSub(&ws.Pm[qe - 10], 1u); }
  }
  __syncwarp();
  {
    uint32_t carry = 0;
    #pragma unroll 1
    for (int base = 0; base < len; base += 32) {
      const int i = base + lane;
      uint32_t incl = i < len ? ws.Pm[i] : 0;
      #pragma unroll
      for (int d = 1; d < 32; d <<= 1) { const uint32_t t = __shfl_up_sync(FULL, incl, d); if (lane >= d) incl += t; }
      incl += carry;
      if (i < len) ws.Pm[i] = incl;
      carry = __shfl_sync(FULL, incl, 31);
    }
  }
  __syncwarp();
  int x = 0;
  #pragma unroll 1
  while (x < len) {
    if (ws.codes[x] > 3) { x++; continue; }
    if (x + 11 > len) break;
    {
      const bool isn = lane < 10 && ws.codes[x + 1 + lane] > 3;
      const uint32_t bal = __ballot_sync(FULL, isn);
      if (bal) { x = x + 1 + (__ffs(bal) - 1) + 1; continue; }
    }
    const int c11 = (int)ws.Pm[x];
    if (c11 == 0) { x += 11; continue; }                     // an empty tile
    int e = x + 11;
    if (c11 >= 20) {
      int c = 0, e20 = -1;
      #pragma unroll 1
      for (int j = 0; j < n; j += 32) {
        const int k = j + lane < n ? (int)sc.ord[j + lane] : -1;
        const uint32_t q = k >= 0 ? sc.runs_q[k] : 0u;
        const int qs = q & 0xffff, qe = q >> 16;
        const unsigned bal = __ballot_sync(FULL, k >= 0 && qs <= x && qe >= x + 11);
        const int cc = __popc(bal);
        if (c + cc >= 20) { e20 = __shfl_sync(FULL, qe, (int)__fns(bal, 0, 20 - c)); break; }
        c += cc;
      }
      if (e20 >= len) break;                                 // the tile would run to the end of the read: stop
      if (ws.codes[e20] > 3) { x = e20 + 1; continue; }     // the base that would extend the tile is an N
      e = e20 + 1;
    }
    #pragma unroll 1
    for (int j = 0; j < n; j += 32) {                        // every run that covers the tile (fewer than 20)
      const int k = j + lane < n ? (int)sc.ord[j + lane] : -1;
      const uint32_t q = k >= 0 ? sc.runs_q[k] : 0u;
      const int qs = q & 0xffff, qe = q >> 16;
      if (__shfl_sync(FULL, qe, 0) < e) break;
      if (k >= 0 && qs <= x && qe >= e) mark_seed(ref, ws.cand, (int)sc.runs_t[k] + (x - qs), e - x);
    }
    x = e;
  }
  __syncwarp();
  return true;
}

__device__ __forceinline__ void write_out(const RefDev& ref, const WarpScratch& ws, int64_t r, uint32_t* candbits, uint32_t* canonbits,
                                          uint32_t* ntask, int lane) {
  uint32_t nf = 0, nr = 0;
  if (lane < MAX_CW) ws.canon[lane] = 0;
  __syncwarp();
  if (lane < ref.CW) {
    uint32_t b = ws.cand[lane];
    candbits[r * ref.CW + lane] = b;
    while (b) {
      const int c = lane * 32 + __ffs(b) - 1;
      b &= b - 1;
      const int cc = ref.canon[c];
      atomicOr(&ws.canon[cc >> 5], 1u << (cc & 31));
    }
  }
  __syncwarp();
  if (lane < ref.CW) {
    const uint32_t b = ws.canon[lane];
    canonbits[r * ref.CW + lane] = b;
    nf = __popc(b & 0x55555555u); nr = __popc(b & 0xAAAAAAAAu);
  }
  nf = warp_sum(nf); nr = warp_sum(nr);
  if (lane == 0) ntask[r] = ref.direct ? nf + nr : (nf + 1) / 2 + (nr + 1) / 2;
}

// GS = false: scratch in shared memory, all reads; overflowing reads are appended to `redo`.
// GS = true : scratch in global memory, only the reads listed in `redo`.
template <bool GS>
__global__ void __launch_bounds__(SEED_WARPS * 32) seed_kernel(RefDev ref, ReadsDev reads, uint32_t* __restrict__ candbits,
                                                               uint32_t* __restrict__ canonbits, uint32_t* __restrict__ ntask, uint32_t* __restrict__ redo,
                                                               unsigned long long* __restrict__ counters,
                                                               uint32_t* __restrict__ gscratch, int t2_bytes_smem) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ uint64_t bar;
  if (GS && counters[4] == 0) return;                     // no read overflowed the shared run list (the usual case): skip the text staging too
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int ml = (reads.max_len + 31) & ~31;
  const size_t wbytes = (warp_scratch_bytes(ml) + 15) & ~(size_t)15;
  const WarpScratch ws = carve(smem_raw + wbytes * warp, ml);
  const uint32_t* T2 = ref.T2;
  if (t2_bytes_smem > 0) {
    uint32_t* t2s = reinterpret_cast<uint32_t*>(smem_raw + wbytes * SEED_WARPS);
    stage_to_smem(t2s, ref.T2, (uint32_t)t2_bytes_smem, &bar);
    T2 = t2s;
  }
  const int64_t gw = (int64_t)blockIdx.x * SEED_WARPS + warp;
  Scratch sc;
  if (GS) {
    uint32_t* base = gscratch + (size_t)gw * (RUNS_GLOBAL * 3);
    sc.runs_q = base; sc.runs_t = base + RUNS_GLOBAL; sc.cov = reinterpret_cast<uint16_t*>(base + 2 * RUNS_GLOBAL);
    sc.ord = reinterpret_cast<uint16_t*>(base + 2 * RUNS_GLOBAL + RUNS_GLOBAL / 2); sc.cap = RUNS_GLOBAL;
  } else {
    sc.runs_q = ws.runs_q; sc.runs_t = ws.runs_t; sc.cov = ws.cov; sc.ord = ws.ord; sc.cap = RUNS_SMEM;
  }
  // reads are handed out SEED_GRAB at a time from a device-wide cursor: the work per read varies several-fold and a CTA may start
  // late (the other reference set's kernels share the SMs), so a static split would end on its unluckiest warp
  const int64_t total = GS ? (int64_t)counters[4] : reads.n;
  unsigned long long* cursor = counters + (GS ? CTR_WORK_REDO : CTR_WORK_SEED);
  int64_t it = 0, it_end = 0;
  for (;;) {
    if (it == it_end) {
      unsigned long long b = 0;
      if (lane == 0) b = atomicAdd(cursor, (unsigned long long)SEED_GRAB);
      it = (int64_t)__shfl_sync(FULL, b, 0);
      if (it >= total) break;
      it_end = min(it + SEED_GRAB, total);
    }
    const int64_t r = GS ? (int64_t)redo[it] : it;
    it++;
    const bool ok = seed_one_read(ref, T2, reads, r, ws, sc, lane);
    if (ok) {
      write_out(ref, ws, r, candbits, canonbits, ntask, lane);
    } else if (!GS) {
      if (lane == 0) { const unsigned long long k = atomicAdd(&counters[4], 1ull); redo[k] = (uint32_t)r; ntask[r] = 0; }
    } else {
      if (lane == 0) { atomicAdd(&counters[6], 1ull); ntask[r] = 0; }    // hard overflow: reported to the host as an error
      if (lane < ref.CW) { candbits[r * ref.CW + lane] = 0; canonbits[r * ref.CW + lane] = 0; }
    }
    __syncwarp();
  }
}

}  // namespace

int launch_seed(const RefDev& ref, const ReadsDev& reads, AlignBuffers& ab, int sm_count, cudaStream_t st, int* launches) {
  const int t2_bytes = t2_words(ref.L) * 4;
  const int t2_smem = t2_bytes <= T2_SMEM_MAX_BYTES ? t2_bytes : 0;
  const int ml = (reads.max_len + 31) & ~31;
  const size_t smem = ((warp_scratch_bytes(ml) + 15) & ~(size_t)15) * SEED_WARPS + t2_smem;
  static std::atomic<unsigned long long> optin0{0}, optin1{0};
  int rc;
  if ((rc = smem_optin(seed_kernel<false>, 200 * 1024, optin0)) || (rc = smem_optin(seed_kernel<true>, 200 * 1024, optin1))) return rc;
  int per_sm = 0;
  CATT_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, seed_kernel<false>, SEED_WARPS * 32, smem));
  if (per_sm < 1) per_sm = 1;
  const int64_t want = (reads.n + SEED_WARPS - 1) / SEED_WARPS;
  int grid = (int)std::min<int64_t>((int64_t)sm_count * per_sm, std::max<int64_t>(want, 1));
  seed_kernel<false><<<grid, SEED_WARPS * 32, smem, st>>>(ref, reads, ab.candbits, ab.canonbits, ab.ntask, ab.gapped_list, ab.counters, nullptr, t2_smem);
  CATT_CUDA(cudaGetLastError());
  // overflow pass (usually an empty list): global scratch sized for sm_count*2 CTAs
  const int grid2 = sm_count * 2;
  const size_t need = (size_t)grid2 * SEED_WARPS * (RUNS_GLOBAL * 3);
  if (ab.seed_scratch_runs < need) {
    if (ab.seed_scratch) cudaFree(ab.seed_scratch);
    CATT_CUDA(cudaMalloc(&ab.seed_scratch, need * sizeof(uint32_t)));
    ab.seed_scratch_runs = need;
  }
  seed_kernel<true><<<grid2, SEED_WARPS * 32, smem, st>>>(ref, reads, ab.candbits, ab.canonbits, ab.ntask, ab.gapped_list, ab.counters, ab.seed_scratch, t2_smem);
  CATT_CUDA(cudaGetLastError());
  *launches += 2;
  return CATT_OK;
}

}  // namespace catt
