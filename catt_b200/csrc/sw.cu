// sw.cu — K2/K3: affine-gap local alignment of each read against its candidate V/J records, hit selection and the
// CIGAR facts the Julia side reads.
//
// Replaces the extension + output half of `bwa mem -A1 -B2 -O6 -E1 -L0 -T12` and the SAM fields consumed through
// Jtool.jl:76-124 / catt.jl:69-132:
//   score     = Smith-Waterman local, match +1, mismatch -2, gap(k) = -(6+k), N = -1, of the strand-normalised read
//               against ONE record; end cell = first maximum in row-major order (read row, then record column);
//   selection = best score over the candidates; ties broken by bwa's hash_64(read ordinal + rank in T order);
//   facts     = soft clip (GetStartCigar), end of the first M block (GetEndCigar), POS, NM, sum(M+I).
//
// Two DP kernels share one layout: G lanes cooperate on one task, each lane owns C record columns and keeps H/E for them in
// registers as packed int16x2 (two alignments per register), rows advance as a skewed wavefront, the last column's H and the
// running F cross to the next lane by warp shuffle, the recurrences are DPX packed add-max / 3-way max instructions
// (VIADDMNMX / VIMNMX3 .S16x2) on +2-biased values (see sw_row).  No tensor cores: there is no dense contraction here.
//   sw_bulk2_kernel  - scores of EVERY candidate: task = one read against two candidate records of the same strand; substitution
//                      scores from a shared-memory query profile (LDS.128), two read rows per step.  The roofline kernel.
//   sw_score_kernel  - TRACK pass, the winner of every mapped read only: task = two reads (one per half), scores by PRMT from
//                      per-column byte tables, plus the first cell that reaches the maximum (end of the alignment).
#include <stdlib.h>

#include <cub/device/device_scan.cuh>

#include "internal.cuh"
#include "tma.cuh"

namespace catt {
namespace {

constexpr unsigned FULL = 0xffffffffu;
constexpr int SW_THREADS = 128;
constexpr uint32_t NEG1 = 0xFFFFFFFFu;   // (-1,-1)
constexpr uint32_t NEG7 = 0xFFF9FFF9u;   // (-7,-7): first gap base = open 6 + extend 1
constexpr uint32_t MINV = 0x80008000u;   // (-32768,-32768)
constexpr uint32_t TWO2 = 0x00020002u;   // the +2 bias of every DP quantity
constexpr int BULK_RUN = 4;              // consecutive tasks per lane group and work-cursor grab in the bulk kernel

__device__ __forceinline__ uint64_t hash_64(uint64_t key) {   // bwa's tie-break hash
  key += ~(key << 32);
  key ^= (key >> 22);
  key += ~(key << 13);
  key ^= (key >> 8);
  key += (key << 3);
  key ^= (key >> 15);
  key += ~(key << 27);
  key ^= (key >> 31);
  return key;
}

// prmt.b32 in its generic form: selector nibble = byte index (0..7) | 8 to replicate that byte's sign bit instead.
// (__byte_perm() only honours 3 bits per nibble, so the sign-replicating form needs the raw instruction.)
__device__ __forceinline__ uint32_t prmt(uint32_t a, uint32_t b, uint32_t sel) {
  uint32_t d;
  asm("prmt.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(sel));
  return d;
}

// Work distribution of the persistent DP kernels: a warp takes its next `count` tasks from a device-wide cursor.  (A static
// partition made every launch as slow as its last-started CTA: the V and J reference sets run on two streams, and a CTA that has
// to wait for a slot held by the other stream's kernel would still own a full 1/grid share of the tasks.)
__device__ __forceinline__ int64_t next_tasks(unsigned long long* work, int count, int lane) {
  unsigned long long b = 0;
  if (lane == 0) b = atomicAdd(work, (unsigned long long)count);
  return (int64_t)__shfl_sync(FULL, b, 0);
}

// SAM-oriented base of a packed read: strand 1 = reverse complement
__device__ __forceinline__ uint32_t read_code(const uint32_t* __restrict__ w, const uint32_t* __restrict__ nm, int len, int i, int strand) {
  const int p = strand ? len - 1 - i : i;
  if ((nm[p >> 5] >> (p & 31)) & 1u) return 4u;
  const uint32_t c = (w[p >> 4] >> ((p & 15) * 2)) & 3u;
  return strand ? 3u - c : c;
}

// ---------------------------------------------------------------------------------------------------------------
// task expansion: canonical-candidate bitmap -> (read, candA, candB) tasks, pairing candidates of the same strand.
// Byte-identical records (hs TRBV: 47 of 168 records in 22 groups) are scored once, through their first copy.
// ---------------------------------------------------------------------------------------------------------------
__global__ void expand_kernel(RefDev ref, int64_t n, const uint32_t* __restrict__ canonbits, const uint32_t* __restrict__ task_off,
                              SwTask* __restrict__ tasks, int64_t cap_tasks, unsigned long long* __restrict__ counters) {
  const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n) return;
  if ((int64_t)task_off[n] > cap_tasks) {                 // task buffer too small: the host grows it and re-runs the stage
    if (r == 0) counters[9] = task_off[n];
    return;
  }
  uint32_t o = task_off[r];
  if (task_off[r + 1] == o) return;
  for (int strand = 0; strand < 2; strand++) {
    int pend = -1;
    for (int w = 0; w < ref.CW; w++) {
      uint32_t b = canonbits[r * ref.CW + w] & (strand ? 0xAAAAAAAAu : 0x55555555u);
      while (b) {
        const int bit = __ffs(b) - 1;
        b &= b - 1;
        const int c = w * 32 + bit;
        if (ref.direct) tasks[o++] = SwTask{(uint32_t)r, (uint16_t)c, (uint16_t)0xFFFF};
        else if (pend < 0) pend = c;
        else { tasks[o++] = SwTask{(uint32_t)r, (uint16_t)pend, (uint16_t)c}; pend = -1; }
      }
    }
    if (pend >= 0) tasks[o++] = SwTask{(uint32_t)r, (uint16_t)pend, (uint16_t)0xFFFF};
  }
}

// ---------------------------------------------------------------------------------------------------------------
// the DP kernel
// ---------------------------------------------------------------------------------------------------------------
// One DP row over this lane's C columns, two candidates per register (int16x2).  Everything is kept biased by +2
// (H2 = H+2, E2 = max(E,0)+2, F2 = max(F,0)+2; flooring E and F at 0 never changes H = max(0, ., E, F)):
//   * the diagonal add t = H2(i-1,j-1) + s becomes a plain 32-bit add, because the low half H2 + s >= 0 never borrows and
//     its carry (present exactly when s_A < 0) is pre-compensated in candidate B's byte table - so the add can issue on
//     the FMA pipe (IMAD) while the ALU pipe, the bottleneck of this kernel, keeps the DPX min/max work;
//   * H2 = max3(t, E2, F2) needs no separate zero floor.
// MODE 0: bulk pass, both halves driven by the same read base (the carry compensation in tabB assumes that).
// MODE 1: TRACK pass (two different reads): plain tables, packed 16x2 add.
// MODE 2: a half whose read base is N (or whose read has ended) scores -1 against everything (rare rows; in the bulk
//         pass both halves at once, keep = 0).
template <int C, int MODE>
__device__ __forceinline__ void sw_row(uint32_t (&Hp)[C], uint32_t (&E)[C], const uint32_t (&tabA)[C], const uint32_t (&tabB)[C],
                                       uint32_t sel, uint32_t keep, uint32_t neg7, uint32_t& hd, uint32_t& f, uint32_t& rowbest) {
  #pragma unroll
  for (int c = 0; c < C; c++) {
    const uint32_t s = prmt(tabA[c], tabB[c], sel);
    uint32_t t;
    if (MODE == 0) t = hd + s;                               // 32-bit add == packed add here (see above)
    else if (MODE == 1) t = __vadd2(hd, s);
    else t = __vadd2(hd, (s & keep) | ~keep);
    const uint32_t h = __vimax3_s16x2(t, E[c], f);          // H2 = max(H(i-1,j-1)+s, E, F, 0) + 2
    hd = Hp[c];
    Hp[c] = h;
    const uint32_t hm7 = __viaddmax_s16x2(h, neg7, TWO2);   // max(H-7, 0) + 2
    E[c] = __viaddmax_s16x2(E[c], NEG1, hm7);           // E(i+1,j)
    f = __viaddmax_s16x2(f, NEG1, hm7);                 // F(i,j+1)
    if (c & 1) rowbest = __vimax3_s16x2(rowbest, Hp[c - 1], h);
    else if (c == C - 1) rowbest = __vimax_s16x2_relu(rowbest, h);
  }
}

__device__ __forceinline__ uint32_t pack_key(int score, int i, int j) {
  return ((uint32_t)score << 20) | ((uint32_t)(1023 - i) << 10) | (uint32_t)(1023 - j);
}

// G lanes per task, C columns per lane; record length <= G*C.
// TRACK = false: scores only (out[t] = scoreA | scoreB << 16) - the bulk pass over every candidate; a task is one read
//                against two of its candidates.
// TRACK = true : also the first cell reaching the maximum in row-major order - run on the one winning candidate per read,
//                which is all the SAM record needs.  `tasks` then holds one single-candidate entry per winner and a
//                task pairs the winners 2t and 2t+1 (two different reads, each driving its own half); out[k] = pack_key
//                of winner k.
// The task count is read on the device (*ntasks_dev) so the stage needs no host round trip.
// Per row the shared array holds the PRMT selector (low 16 bits) and one "scores -1" flag per half (bits 16, 17).
template <int G, int C, bool TRACK>
__global__ void __launch_bounds__(SW_THREADS) sw_score_kernel(RefDev ref, ReadsDev reads, const SwTask* __restrict__ tasks,
                                                              const uint32_t* __restrict__ ntasks_dev, int64_t ntasks_cap,
                                                              uint32_t* __restrict__ out, int tb_bytes_smem, int ml,
                                                              unsigned long long* __restrict__ work) {
  constexpr int GPW = 32 / G;                         // tasks per warp
  constexpr int GPB = GPW * (SW_THREADS / 32);        // tasks per block
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ uint64_t bar;
  const int64_t nent = (int64_t)*ntasks_dev;          // entries in `tasks`
  if (nent > ntasks_cap) return;                      // expand_kernel flagged the overflow; nothing valid to score
  const int64_t ntasks = TRACK ? (nent + 1) >> 1 : nent;
  if ((int64_t)blockIdx.x * GPB >= ntasks) return;
  uint32_t* sel_all = reinterpret_cast<uint32_t*>(smem_raw);              // [GPB][ml]
  const uint8_t* Tb = ref.Tb;
  if (tb_bytes_smem > 0) {
    uint8_t* tbs = smem_raw + (size_t)GPB * ml * sizeof(uint32_t);
    stage_to_smem(tbs, ref.Tb, (uint32_t)tb_bytes_smem, &bar);
    Tb = tbs;
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int g = lane % G, grp = lane / G;
  uint32_t* sel_arr = sel_all + (size_t)(warp * GPW + grp) * ml;

  for (;;) {
    const int64_t base = next_tasks(work, GPW, lane);
    if (base >= ntasks) break;
    const int64_t t = base + grp;
    const bool live = t < ntasks;
    int m = 0;
    uint32_t tabA[C], tabB[C], Hp[C], E[C];
    #pragma unroll
    for (int c = 0; c < C; c++) { tabA[c] = 0xFEFEFEFEu; tabB[c] = TRACK ? 0xFEFEFEFEu : 0xFDFDFDFDu; Hp[c] = TWO2; E[c] = TWO2; }
    if (live) {
      int ca, cb = 0xFFFF;
      if (!TRACK) {
        const SwTask tk = tasks[t];
        ca = tk.a; cb = tk.b;
        const int strand = ca & 1;
        m = reads.len[tk.read];
        const uint32_t* __restrict__ w = reads.words + reads.off[tk.read];
        const uint32_t* __restrict__ nm = w + read_base_words(m);
        for (int i = g; i < m; i += G) {
          const uint32_t q = read_code(w, nm, m, i, strand);
          sel_arr[i] = q > 3 ? (0x30000u | 0xC480u) : 0xC480u + 0x1111u * q;   // bytes: {A[q], sign(A[q]), B[q], sign(B[q])}
        }
      } else {
        const SwTask ta = tasks[2 * t];
        const bool hasb = 2 * t + 1 < nent;
        const SwTask tb2 = hasb ? tasks[2 * t + 1] : ta;
        ca = ta.a; cb = hasb ? tb2.a : 0xFFFF;
        const int ma = reads.len[ta.read], mb = hasb ? reads.len[tb2.read] : 0;
        m = max(ma, mb);
        const uint32_t* __restrict__ wa = reads.words + reads.off[ta.read];
        const uint32_t* __restrict__ na = wa + read_base_words(ma);
        const uint32_t* __restrict__ wb = reads.words + reads.off[tb2.read];
        const uint32_t* __restrict__ nb = wb + read_base_words(reads.len[tb2.read]);
        for (int i = g; i < m; i += G) {
          const uint32_t qa = i < ma ? read_code(wa, na, ma, i, ca & 1) : 4u;
          const uint32_t qb = i < mb ? read_code(wb, nb, mb, i, cb & 1) : 4u;
          sel_arr[i] = (0x80u + 0x11u * (qa & 3u)) | ((0xC4u + 0x11u * (qb & 3u)) << 8) | (qa > 3 ? 0x10000u : 0u) | (qb > 3 ? 0x20000u : 0u);
        }
      }
      // per-column byte tables: byte q of tabA = score of read base q against candidate A's base in this column (+1 / -2);
      // byte q of tabB = candidate B's score minus 1 where A's score is negative (the carry the 32-bit add will produce)
      {
        const int ra = ca >> 1, oa = ref.rec_off[ra], la = ref.rec_len[ra];
        int rb = 0, ob = 0, lb = 0;
        if (cb != 0xFFFF) { rb = cb >> 1; ob = ref.rec_off[rb]; lb = ref.rec_len[rb]; }
        #pragma unroll
        for (int c = 0; c < C; c++) {
          const int j = g * C + c;
          uint32_t amask = 0;                                                    // 0xFF in the byte of A's base
          if (j < la) { amask = 0xFFu << (8 * Tb[oa + j]); tabA[c] = 0xFEFEFEFEu ^ amask; }
          uint32_t tb = 0xFEFEFEFEu;
          if (j < lb) tb ^= 0xFFu << (8 * Tb[ob + j]);
          tabB[c] = TRACK ? tb : tb - (0x01010101u & ~amask);
        }
      }
    }
    __syncwarp();
    int steps = live ? m + G - 1 : 0;
    #pragma unroll
    for (int d = 16; d; d >>= 1) steps = max(steps, __shfl_xor_sync(FULL, steps, d));

    uint32_t lastH = TWO2, lastF = TWO2, diag_in = TWO2, best = TWO2;
    uint32_t neg7;                                          // kept in a register: as an immediate ptxas re-materialises it per cell
    neg7 = NEG7 ^ (uint32_t)((unsigned long long)ntasks_cap >> 62);   // == NEG7; opaque to constant propagation
    const uint32_t sel_base = smem_u32(sel_arr);            // shared-window address, so the row load is one LDS off a running index
    int bi_lo = 0, bj_lo = 0, bi_hi = 0, bj_hi = 0;
    for (int st = 0; st < steps; st++) {
      uint32_t inH = __shfl_up_sync(FULL, lastH, 1, G);
      uint32_t inF = __shfl_up_sync(FULL, lastF, 1, G);
      if (g == 0) { inH = TWO2; inF = TWO2; }
      const int i = st - g;
      if (i >= 0 && i < m) {
        uint32_t selw;
        asm volatile("ld.shared.u32 %0, [%1];" : "=r"(selw) : "r"(sel_base + 4u * (uint32_t)i));
        const uint32_t sel = selw & 0xffffu, fl = selw >> 16;
        uint32_t hd = diag_in, f = inF;       // F(i, first column) arrives already advanced from the left neighbour
        diag_in = inH;
        if (!TRACK) {
          if (fl) sw_row<C, 2>(Hp, E, tabA, tabB, sel, 0u, neg7, hd, f, best);
          else sw_row<C, 0>(Hp, E, tabA, tabB, sel, 0u, neg7, hd, f, best);
          lastH = Hp[C - 1];
          lastF = f;
        } else {
          uint32_t rowbest = TWO2;
          if (fl) sw_row<C, 2>(Hp, E, tabA, tabB, sel, ((fl & 1u) ? 0u : 0xFFFFu) | ((fl & 2u) ? 0u : 0xFFFF0000u), neg7, hd, f, rowbest);
          else sw_row<C, 1>(Hp, E, tabA, tabB, sel, 0u, neg7, hd, f, rowbest);
          lastH = Hp[C - 1];
          lastF = f;
          const uint32_t nb = __vimax_s16x2_relu(best, rowbest);
          if (nb != best) {                                 // a half improved: record the first cell reaching it
            const int rl = (int)(short)(rowbest & 0xffff), bl = (int)(short)(best & 0xffff);
            const int rh = (int)(short)(rowbest >> 16), bh = (int)(short)(best >> 16);
            if (rl > bl) {
              bi_lo = i;
              int jj = 0;
              #pragma unroll
              for (int c = C - 1; c >= 0; c--) if ((int)(short)(Hp[c] & 0xffff) == rl) jj = c;
              bj_lo = g * C + jj;
            }
            if (rh > bh) {
              bi_hi = i;
              int jj = 0;
              #pragma unroll
              for (int c = C - 1; c >= 0; c--) if ((int)(short)(Hp[c] >> 16) == rh) jj = c;
              bj_hi = g * C + jj;
            }
            best = nb;
          }
        }
      }
    }
    if (!TRACK) {
      #pragma unroll
      for (int d = G / 2; d; d >>= 1) best = __vimax_s16x2_relu(best, __shfl_xor_sync(FULL, best, d, G));
      if (live && g == 0) out[t] = best - TWO2;
    } else {
      // reduce over the G lanes of the task: max score, then earliest row, then earliest column
      uint32_t klo = pack_key((int)(short)(best & 0xffff) - 2, bi_lo, bj_lo);
      uint32_t khi = pack_key((int)(short)(best >> 16) - 2, bi_hi, bj_hi);
      #pragma unroll
      for (int d = G / 2; d; d >>= 1) {
        klo = max(klo, __shfl_xor_sync(FULL, klo, d, G));
        khi = max(khi, __shfl_xor_sync(FULL, khi, d, G));
      }
      if (live && g == 0) { out[2 * t] = klo; if (2 * t + 1 < nent) out[2 * t + 1] = khi; }
    }
    __syncwarp();
  }
}

template <int G, int C, bool TRACK>
int launch_sw_t(const RefDev& ref, const ReadsDev& reads, const SwTask* tasks, const uint32_t* ntasks_dev, int64_t cap, uint32_t* res,
                unsigned long long* work, int sm_count, cudaStream_t st) {
  constexpr int GPB = (32 / G) * (SW_THREADS / 32);
  const int ml = (reads.max_len + 3) & ~3;
  const int tb_bytes = (ref.L + 64) & ~15;
  const size_t sel_bytes = (size_t)GPB * ml * sizeof(uint32_t);
  const int tb_smem = (sel_bytes + tb_bytes <= 100 * 1024) ? tb_bytes : 0;
  const size_t smem = sel_bytes + tb_smem;
  auto k = sw_score_kernel<G, C, TRACK>;
  static std::atomic<unsigned long long> optin{0};
  int per_sm = 0;
  { const int rc = smem_optin(k, 200 * 1024, optin); if (rc) return rc; }
  CATT_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k, SW_THREADS, smem));
  if (per_sm < 1) per_sm = 1;
  const int64_t ntasks_max = TRACK ? (cap + 1) / 2 : cap;
  const int64_t want = (ntasks_max + GPB - 1) / GPB;
  const int grid = (int)std::max<int64_t>(1, std::min<int64_t>((int64_t)sm_count * per_sm, want));
  k<<<grid, SW_THREADS, smem, st>>>(ref, reads, tasks, ntasks_dev, cap, res, tb_smem, ml, work);
  CATT_CUDA(cudaGetLastError());
  return CATT_OK;
}

// ---------------------------------------------------------------------------------------------------------------
// the bulk score pass with a shared-memory query profile.
// Same recurrence, layout and +2 bias as sw_score_kernel (the TRACK pass above), but the substitution scores of a row come from
// a per-task profile in shared memory instead of one PRMT per cell: prof[q][column] = (s_B - [s_A < 0]) << 16 | s_A for read
// base q = 0..3 and q = 4 (N: -1 against everything), laid out so that the G lanes of a task read one contiguous 16-byte word
// each per LDS.128 (conflict free, also across the skewed rows, whose profile rows are 128-byte multiples apart).  That moves 1
// of the 5.5 ALU-pipe instructions per cell pair to the otherwise idle LSU pipe.
// ---------------------------------------------------------------------------------------------------------------
// Shared-memory layout of one task's profile (bulk and grouped TRACK kernels).  A profile row (one read symbol) holds the lane
// groups' 16-byte words chunk by chunk: word (chunk, g) at chunk * G * 16 + g * 16.  With G = 4 two tasks share each 128-byte bank
// phase of an LDS.128, so they must sit in different 64-byte halves WHATEVER rows they read: rows are padded to multiples of 128
// bytes and the task stride is an odd multiple of 64.  (With the natural 448-byte rows of the 4 x 25 tiling the half depended on
// the read symbol: ncu counted 1.2 - 2 wavefronts per ideal one.)  A column count that is not a multiple of 4 keeps its last C % 4
// columns in a small tail [symbol][lane][column] read with scalar loads instead of padding every row by a chunk.
template <int G, int C>
struct ProfLayout {
  static constexpr int CHF = C / 4, TAIL = C % 4;
  static constexpr int ROW_RAW = G * CHF * 16;
  static constexpr int ROWB = G == 4 ? ((ROW_RAW + 127) & ~127) : ROW_RAW;
  static constexpr int TAIL_ROWB = G * TAIL * 4;
  static constexpr int SLOT_RAW = 5 * ROWB + 5 * TAIL_ROWB;
  static constexpr int SLOTB = G == 4 ? ((SLOT_RAW + 63) / 128) * 128 + 64 : ((SLOT_RAW + 15) & ~15);
  // word index (from the profile base, in 32-bit words) of symbol q, lane g, column c
  __device__ static __forceinline__ int word(int q, int g, int c) {
    return c < 4 * CHF ? q * (ROWB / 4) + ((c >> 2) * G + g) * 4 + (c & 3) : 5 * (ROWB / 4) + q * (TAIL_ROWB / 4) + g * TAIL + (c - 4 * CHF);
  }
};
// per-task stride of the read-code arrays, in bytes: a word count of 9 mod 32 puts the eight tasks of a warp nine banks apart
__host__ __device__ inline int qarr_stride(int max_len) {
  int w = (max_len + 1 + 3) / 4;
  w += (9 - w % 32 + 32) % 32;
  return 4 * w;
}

// Two rows per wavefront step: lane g works on read rows 2(st-g) and 2(st-g)+1 for its C columns, column by column, the second
// row one cell behind the first.  Same cells, same operations per cell - but two independent F chains per thread (the ALU pipe
// was stalling on the 3-instruction dependency chain of a single row) and half as many shuffles / loop iterations per row.
template <int G, int C>
__global__ void __launch_bounds__(SW_THREADS) sw_bulk2_kernel(RefDev ref, ReadsDev reads, const SwTask* __restrict__ tasks,
                                                              const uint32_t* __restrict__ ntasks_dev, int64_t ntasks_cap,
                                                              uint32_t* __restrict__ out, int ml, unsigned long long* __restrict__ work) {
  constexpr int GPW = 32 / G;
  constexpr int GPB = GPW * (SW_THREADS / 32);
  using L = ProfLayout<G, C>;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int64_t ntasks = (int64_t)*ntasks_dev;
  if (ntasks > ntasks_cap) return;
  if ((int64_t)blockIdx.x * GPB >= ntasks) return;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int g = lane % G, grp = lane / G;
  const int slot = warp * GPW + grp;
  uint32_t* prof = reinterpret_cast<uint32_t*>(smem_raw + (size_t)slot * L::SLOTB);
  uint8_t* qarr = smem_raw + (size_t)GPB * L::SLOTB + (size_t)slot * ml;       // ml is even and > max_len
  const uint32_t prof_base = smem_u32(prof) + (uint32_t)g * 16u;
  const uint32_t tail_base = smem_u32(prof) + (uint32_t)(5 * L::ROWB + g * L::TAIL * 4);
  const uint8_t* __restrict__ Tb = ref.Tb;
  const uint32_t neg7 = NEG7 ^ (uint32_t)((unsigned long long)ntasks_cap >> 62);   // == NEG7, opaque to constant propagation

  // A warp takes BULK_RUN tasks per lane group at a time and every group works through CONSECUTIVE ones: the tasks of a read are
  // adjacent in the list (about seven per read for hs TRBV), so the group's read-code array is rebuilt only when the read or the
  // strand changes
  int64_t cur_read = -1;
  int cur_strand = -1, cur_m = 0;
  for (;;) {
    const int64_t base0 = next_tasks(work, GPW * BULK_RUN, lane);
    if (base0 >= ntasks) break;
   #pragma unroll 1
   for (int kk = 0; kk < BULK_RUN; kk++) {
    const int64_t t = base0 + (int64_t)grp * BULK_RUN + kk;
    if (base0 + kk >= ntasks) break;                    // (warp-uniform: no group has a task kk any more)
    const bool live = t < ntasks;
    int m = 0;
    if (live) {
      const SwTask tk = tasks[t];
      const int strand = tk.a & 1;
      if ((int64_t)tk.read != cur_read || strand != cur_strand) {
        cur_read = tk.read; cur_strand = strand;
        cur_m = reads.len[tk.read];
        const uint32_t* __restrict__ w = reads.words + reads.off[tk.read];
        const uint32_t* __restrict__ nm = w + read_base_words(cur_m);
        for (int i = g; i < cur_m; i += G) qarr[i] = (uint8_t)read_code(w, nm, cur_m, i, strand);
        if (g == 0 && (cur_m & 1)) qarr[cur_m] = 4;     // the phantom last row of an odd-length read scores -1 everywhere
      }
      m = cur_m;
      const int ra = tk.a >> 1, oa = ref.rec_off[ra], la = ref.rec_len[ra];
      int ob = 0, lb = 0;
      if (tk.b != 0xFFFF) { const int rb = tk.b >> 1; ob = ref.rec_off[rb]; lb = ref.rec_len[rb]; }
      // profile word of read symbol q for a column holding bases (ba, bb); four columns go out as one 16-byte store
      auto pw = [](int q, int ba, int bb) -> uint32_t {
        const int sa = q == ba ? 1 : -2, sb = (q == bb ? 1 : -2) - (sa < 0);
        return ((uint32_t)sb << 16) | ((uint32_t)sa & 0xFFFFu);
      };
      #pragma unroll
      for (int ch = 0; ch < L::CHF; ch++) {
        int ba[4], bb[4];
        #pragma unroll
        for (int k = 0; k < 4; k++) { const int j = g * C + 4 * ch + k; ba[k] = j < la ? (int)Tb[oa + j] : 7; bb[k] = j < lb ? (int)Tb[ob + j] : 7; }
        #pragma unroll
        for (int q = 0; q < 4; q++)
          *reinterpret_cast<uint4*>(prof + L::word(q, g, 4 * ch)) = make_uint4(pw(q, ba[0], bb[0]), pw(q, ba[1], bb[1]), pw(q, ba[2], bb[2]), pw(q, ba[3], bb[3]));
        *reinterpret_cast<uint4*>(prof + L::word(4, g, 4 * ch)) = make_uint4(0xFFFEFFFFu, 0xFFFEFFFFu, 0xFFFEFFFFu, 0xFFFEFFFFu);
      }
      #pragma unroll
      for (int k = 0; k < L::TAIL; k++) {
        const int c = 4 * L::CHF + k, j = g * C + c;
        const int ba = j < la ? (int)Tb[oa + j] : 7, bb = j < lb ? (int)Tb[ob + j] : 7;
        #pragma unroll
        for (int q = 0; q < 4; q++) prof[L::word(q, g, c)] = pw(q, ba, bb);
        prof[L::word(4, g, c)] = 0xFFFEFFFFu;
      }
    }
    __syncwarp();
    const int rows2 = (m + 1) >> 1;
    int steps = live ? rows2 + G - 1 : 0;
    #pragma unroll
    for (int d = 16; d; d >>= 1) steps = max(steps, __shfl_xor_sync(FULL, steps, d));

    uint32_t Hp[C], E[C];
    #pragma unroll
    for (int c = 0; c < C; c++) { Hp[c] = TWO2; E[c] = TWO2; }
    uint32_t lastH0 = TWO2, lastH1 = TWO2, lastF0 = TWO2, lastF1 = TWO2, diag_in = TWO2, best = TWO2;
    const uint32_t q_base = smem_u32(qarr);
    for (int st = 0; st < steps; st++) {
      uint32_t inH0 = __shfl_up_sync(FULL, lastH0, 1, G);
      uint32_t inH1 = __shfl_up_sync(FULL, lastH1, 1, G);
      uint32_t inF0 = __shfl_up_sync(FULL, lastF0, 1, G);
      uint32_t inF1 = __shfl_up_sync(FULL, lastF1, 1, G);
      if (g == 0) { inH0 = TWO2; inH1 = TWO2; inF0 = TWO2; inF1 = TWO2; }
      const int i2 = st - g;
      if (i2 >= 0 && i2 < rows2) {
        uint32_t qq;
        asm volatile("ld.shared.u16 %0, [%1];" : "=r"(qq) : "r"(q_base + 2u * (uint32_t)i2));
        const uint32_t q0 = qq & 0xffu, q1 = qq >> 8;
        const uint32_t row0 = prof_base + q0 * (uint32_t)L::ROWB;
        const uint32_t row1 = prof_base + q1 * (uint32_t)L::ROWB;
        uint32_t hd0 = diag_in, f0 = inF0;    // row 2*i2:   diagonal = H(row-1, left neighbour's last column)
        uint32_t hd1 = inH0, f1 = inF1;       // row 2*i2+1: diagonal = H(row 2*i2, left neighbour's last column)
        diag_in = inH1;
        uint32_t h0 = TWO2;
        auto cell = [&](int c, uint32_t sc0, uint32_t sc1) {
          const uint32_t t0 = hd0 + sc0;                             // packed adds as plain 32-bit adds (see sw_row)
          h0 = __vimax3_s16x2(t0, E[c], f0);
          hd0 = Hp[c];
          const uint32_t hm0 = __viaddmax_s16x2(h0, neg7, TWO2);
          const uint32_t e0 = __viaddmax_s16x2(E[c], NEG1, hm0);
          f0 = __viaddmax_s16x2(f0, NEG1, hm0);
          const uint32_t t1 = hd1 + sc1;
          const uint32_t h1 = __vimax3_s16x2(t1, e0, f1);
          hd1 = h0;
          const uint32_t hm1 = __viaddmax_s16x2(h1, neg7, TWO2);
          E[c] = __viaddmax_s16x2(e0, NEG1, hm1);
          f1 = __viaddmax_s16x2(f1, NEG1, hm1);
          Hp[c] = h1;
          best = __vimax3_s16x2(best, h0, h1);
        };
        #pragma unroll
        for (int ch = 0; ch < L::CHF; ch++) {
          uint32_t s0[4], s1[4];
          asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(s0[0]), "=r"(s0[1]), "=r"(s0[2]), "=r"(s0[3]) : "r"(row0 + (uint32_t)(ch * G * 16)));
          asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(s1[0]), "=r"(s1[1]), "=r"(s1[2]), "=r"(s1[3]) : "r"(row1 + (uint32_t)(ch * G * 16)));
          #pragma unroll
          for (int k = 0; k < 4; k++) cell(4 * ch + k, s0[k], s1[k]);
        }
        #pragma unroll
        for (int k = 0; k < L::TAIL; k++) {
          uint32_t a0, a1;
          asm volatile("ld.shared.u32 %0, [%1];" : "=r"(a0) : "r"(tail_base + q0 * (uint32_t)L::TAIL_ROWB + (uint32_t)(k * 4)));
          asm volatile("ld.shared.u32 %0, [%1];" : "=r"(a1) : "r"(tail_base + q1 * (uint32_t)L::TAIL_ROWB + (uint32_t)(k * 4)));
          cell(4 * L::CHF + k, a0, a1);
        }
        lastH0 = h0; lastH1 = Hp[C - 1];
        lastF0 = f0; lastF1 = f1;
      }
    }
    #pragma unroll
    for (int d = G / 2; d; d >>= 1) best = __vimax_s16x2_relu(best, __shfl_xor_sync(FULL, best, d, G));
    if (live && g == 0) out[t] = best - TWO2;
    __syncwarp();
   }
  }
}

template <int G, int C>
int launch_sw_bulk(const RefDev& ref, const ReadsDev& reads, const SwTask* tasks, const uint32_t* ntasks_dev, int64_t cap, uint32_t* res,
                   unsigned long long* work, int sm_count, cudaStream_t st) {
  constexpr int GPB = (32 / G) * (SW_THREADS / 32);
  const int ml = qarr_stride(reads.max_len);                     // room for the phantom row of odd-length reads
  const size_t smem = (size_t)GPB * (ProfLayout<G, C>::SLOTB + ml);
  auto k = sw_bulk2_kernel<G, C>;
  static std::atomic<unsigned long long> optin{0};
  int per_sm = 0;
  { const int rc = smem_optin(k, 200 * 1024, optin); if (rc) return rc; }
  CATT_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k, SW_THREADS, smem));
  if (per_sm < 1) per_sm = 1;
  const int64_t want = (cap + GPB - 1) / GPB;
  const int grid = (int)std::max<int64_t>(1, std::min<int64_t>((int64_t)sm_count * per_sm, want));
  k<<<grid, SW_THREADS, smem, st>>>(ref, reads, tasks, ntasks_dev, cap, res, ml, work);
  CATT_CUDA(cudaGetLastError());
  return CATT_OK;
}

int dispatch_sw_bulk(const RefDev& ref, const ReadsDev& reads, const SwTask* tasks, const uint32_t* ntasks_dev, int64_t cap, uint32_t* res,
                     unsigned long long* work, int sm_count, cudaStream_t st) {
  const int n = ref.maxlen;
  if (n <= 32) return launch_sw_bulk<4, 8>(ref, reads, tasks, ntasks_dev, cap, res, work, sm_count, st);
  if (n <= 48) return launch_sw_bulk<4, 12>(ref, reads, tasks, ntasks_dev, cap, res, work, sm_count, st);
  if (n <= 64) return launch_sw_bulk<8, 8>(ref, reads, tasks, ntasks_dev, cap, res, work, sm_count, st);
  if (n <= 100) return launch_sw_bulk<4, 25>(ref, reads, tasks, ntasks_dev, cap, res, work, sm_count, st);
  if (n <= 104) return launch_sw_bulk<8, 13>(ref, reads, tasks, ntasks_dev, cap, res, work, sm_count, st);
  // long records (IGHV, pig TRBV: 280-320 nt): few lanes with many columns keep the skew short (G - 1 idle steps per task);
  // 40 columns per lane cost 152 registers and 2 CTAs / SM, which the two-row kernel tolerates (IGH: 4485 -> 5077 GCUPS vs 16 x 20)
  if (n <= 160) return launch_sw_bulk<4, 40>(ref, reads, tasks, ntasks_dev, cap, res, work, sm_count, st);
  if (n <= 320) return launch_sw_bulk<8, 40>(ref, reads, tasks, ntasks_dev, cap, res, work, sm_count, st);
  set_error("record length %d exceeds the SW kernel's tiling", n);
  return CATT_E_LIMIT;
}

// ---------------------------------------------------------------------------------------------------------------
// TRACK pass, profile version: the end cell of every winner whose score is already known (bulk pass + selection).
// The winners are grouped by record first (counting sort, bins padded to even sizes), so a task = two reads against the SAME
// record and the G lanes keep ONE record profile in shared memory - prof[q][column] = +1 / -2 as a 32-bit integer, rebuilt
// only when the group moves on to the next record (a bin holds thousands of winners).  Each half looks its own read base up:
//   t = H(i-1,j-1) + prof[q_lo][j] + (prof[q_hi][j] << 16)
// (the sign-extended low score borrows from nothing above bit 15 because the biased H is >= 2: the same carry argument as
// sw_row, resolved by integer multiply-add on the FMA pipe).  Two read rows per step like sw_bulk2_kernel.  Since the maximum S is
// known, "the first cell reaching the maximum" is the first cell that EQUALS S: one packed compare per row; the column search
// runs once per read and lane.  out[k] = pack_key(S, i, j) of winner k - what finish_kernel reads.
// ---------------------------------------------------------------------------------------------------------------
constexpr uint32_t NO_WINNER = 0xFFFFFFFFu;

__global__ void whist_kernel(int64_t n, const SwTask* __restrict__ wtasks, const unsigned long long* __restrict__ counters, uint32_t* __restrict__ hist) {
  const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const bool live = k < n && k < (int64_t)counters[7];
  const unsigned act = __ballot_sync(FULL, live);
  if (!live) return;
  const int bin = wtasks[k].a >> 1;
  const unsigned peers = __match_any_sync(act, bin);
  if ((int)(threadIdx.x & 31) == __ffs(peers) - 1) atomicAdd(&hist[bin], (uint32_t)__popc(peers));
}

// one block: bins -> even-padded offsets; cursor[bin] = first slot, the odd bins' last slot is marked empty, *nslots = total
__global__ void wscan_kernel(int nrec, const uint32_t* __restrict__ hist, uint32_t* __restrict__ cursor, uint32_t* __restrict__ order,
                             uint32_t* __restrict__ nslots) {
  __shared__ uint32_t sh[MAX_REC + 1];
  const int b = threadIdx.x;
  const uint32_t c = b < nrec ? hist[b] : 0u;
  sh[b] = (c + 1u) & ~1u;
  __syncthreads();
  if (b == 0) { uint32_t run = 0; for (int i = 0; i < nrec; i++) { const uint32_t v = sh[i]; sh[i] = run; run += v; } sh[nrec] = run; }
  __syncthreads();
  if (b < nrec) { cursor[b] = sh[b]; if (c & 1u) order[sh[b] + c] = NO_WINNER; }
  if (b == 0) *nslots = sh[nrec];
}

__global__ void wscatter_kernel(int64_t n, const SwTask* __restrict__ wtasks, const unsigned long long* __restrict__ counters, uint32_t* __restrict__ cursor,
                                uint32_t* __restrict__ order) {
  const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const bool live = k < n && k < (int64_t)counters[7];
  const unsigned act = __ballot_sync(FULL, live);
  if (!live) return;
  const int bin = wtasks[k].a >> 1, lane = threadIdx.x & 31;
  const unsigned peers = __match_any_sync(act, bin);
  const int leader = __ffs(peers) - 1;
  uint32_t base = 0;
  if (lane == leader) base = atomicAdd(&cursor[bin], (uint32_t)__popc(peers));
  base = __shfl_sync(peers, base, leader);
  order[base + __popc(peers & ((1u << lane) - 1u))] = (uint32_t)k;
}

template <int G, int C>
__global__ void __launch_bounds__(SW_THREADS) sw_track2_kernel(RefDev ref, ReadsDev reads, const SwTask* __restrict__ wtasks,
                                                               const uint32_t* __restrict__ order, const uint32_t* __restrict__ nslots_dev,
                                                               uint32_t* __restrict__ out, int ml, unsigned long long* __restrict__ work) {
  constexpr int GPW = 32 / G;
  constexpr int GPB = GPW * (SW_THREADS / 32);
  using L = ProfLayout<G, C>;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int64_t ntasks = (int64_t)(*nslots_dev >> 1);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int g = lane % G, grp = lane / G;
  const int slot = warp * GPW + grp;
  int32_t* prof = reinterpret_cast<int32_t*>(smem_raw + (size_t)slot * L::SLOTB);
  uint16_t* qarr = reinterpret_cast<uint16_t*>(smem_raw + (size_t)GPB * L::SLOTB + (size_t)slot * ml);   // ml bytes: 4 x 3-bit read codes per row pair
  const uint32_t prof_base = smem_u32(prof) + (uint32_t)g * 16u;
  const uint32_t tail_base = smem_u32(prof) + (uint32_t)(5 * L::ROWB + g * L::TAIL * 4);
  const uint8_t* __restrict__ Tb = ref.Tb;
  const uint32_t neg7 = NEG7 ^ (uint32_t)((unsigned long long)ml >> 30);        // == NEG7, opaque to constant propagation
  int cur_rid = -1;

  for (;;) {
    const int64_t base = next_tasks(work, GPW, lane);
    if (base >= ntasks) break;
    const int64_t t = base + grp;
    const bool live = t < ntasks;
    int m = 0;
    uint32_t ka = NO_WINNER, kb = NO_WINNER, tgt = 0x7FFF7FFFu;
    if (live) {
      ka = order[2 * t]; kb = order[2 * t + 1];
      const SwTask ta = wtasks[ka];
      const SwTask tb = kb != NO_WINNER ? wtasks[kb] : ta;
      const int rid = ta.a >> 1;
      if (rid != cur_rid) {                                 // the group's record profile (rare: winners are grouped by record)
        cur_rid = rid;
        const int oa = ref.rec_off[rid], la = ref.rec_len[rid];
        #pragma unroll
        for (int c = 0; c < C; c++) {
          const int j = g * C + c;
          const int ba = j < la ? (int)Tb[oa + j] : 7;
          #pragma unroll
          for (int q = 0; q < 4; q++) prof[L::word(q, g, c)] = q == ba ? 1 : -2;
          prof[L::word(4, g, c)] = -1;
        }
      }
      const int ma = reads.len[ta.read], mb = kb != NO_WINNER ? reads.len[tb.read] : 0;
      m = max(ma, mb);
      const uint32_t* __restrict__ wa = reads.words + reads.off[ta.read];
      const uint32_t* __restrict__ na = wa + read_base_words(ma);
      const uint32_t* __restrict__ wb = reads.words + reads.off[tb.read];
      const uint32_t* __restrict__ nb = wb + read_base_words(reads.len[tb.read]);
      const int sa = ta.a & 1, sb = tb.a & 1;
      const int rows2 = (m + 1) >> 1;
      for (int i2 = g; i2 < rows2; i2 += G) {                // rows past a read's end score -1 everywhere (they cannot reach S)
        const int i = 2 * i2;
        const uint32_t a0 = i < ma ? read_code(wa, na, ma, i, sa) : 4u, a1 = i + 1 < ma ? read_code(wa, na, ma, i + 1, sa) : 4u;
        const uint32_t b0 = i < mb ? read_code(wb, nb, mb, i, sb) : 4u, b1 = i + 1 < mb ? read_code(wb, nb, mb, i + 1, sb) : 4u;
        qarr[i2] = (uint16_t)(a0 | (a1 << 3) | (b0 << 6) | (b1 << 9));
      }
      tgt = ((uint32_t)ta.b + 2u) | ((kb != NO_WINNER ? (uint32_t)tb.b + 2u : 0x7FFFu) << 16);
    }
    __syncwarp();
    const int rows2 = (m + 1) >> 1;
    int steps = live ? rows2 + G - 1 : 0;
    #pragma unroll
    for (int d = 16; d; d >>= 1) steps = max(steps, __shfl_xor_sync(FULL, steps, d));

    uint32_t Hp[C], E[C];
    #pragma unroll
    for (int c = 0; c < C; c++) { Hp[c] = TWO2; E[c] = TWO2; }
    uint32_t lastH0 = TWO2, lastH1 = TWO2, lastF0 = TWO2, lastF1 = TWO2, diag_in = TWO2;
    uint32_t key_lo = 0, key_hi = 0;                        // (1023 - i) << 10 | (1023 - j) of this lane's first cell equal to S
    const uint32_t q_base = smem_u32(qarr);
    for (int st = 0; st < steps; st++) {
      uint32_t inH0 = __shfl_up_sync(FULL, lastH0, 1, G);
      uint32_t inH1 = __shfl_up_sync(FULL, lastH1, 1, G);
      uint32_t inF0 = __shfl_up_sync(FULL, lastF0, 1, G);
      uint32_t inF1 = __shfl_up_sync(FULL, lastF1, 1, G);
      if (g == 0) { inH0 = TWO2; inH1 = TWO2; inF0 = TWO2; inF1 = TWO2; }
      const int i2 = st - g;
      if (i2 >= 0 && i2 < rows2) {
        uint32_t qq;
        asm volatile("ld.shared.u16 %0, [%1];" : "=r"(qq) : "r"(q_base + 2u * (uint32_t)i2));
        const uint32_t q0l = qq & 7u, q1l = (qq >> 3) & 7u, q0h = (qq >> 6) & 7u, q1h = qq >> 9;
        const uint32_t r0l = prof_base + q0l * (uint32_t)L::ROWB, r1l = prof_base + q1l * (uint32_t)L::ROWB;
        const uint32_t r0h = prof_base + q0h * (uint32_t)L::ROWB, r1h = prof_base + q1h * (uint32_t)L::ROWB;
        uint32_t hd0 = diag_in, f0 = inF0;
        uint32_t hd1 = inH0, f1 = inF1;
        diag_in = inH1;
        uint32_t h0v[C];
        uint32_t rb0 = TWO2, rb1 = TWO2;
        auto cell = [&](int c, int32_t lo0, int32_t hi0, int32_t lo1, int32_t hi1) {
          const uint32_t t0 = (uint32_t)(hi0 * 65536 + (int32_t)(hd0 + (uint32_t)lo0));
          const uint32_t h0 = __vimax3_s16x2(t0, E[c], f0);
          hd0 = Hp[c];
          const uint32_t hm0 = __viaddmax_s16x2(h0, neg7, TWO2);
          const uint32_t e0 = __viaddmax_s16x2(E[c], NEG1, hm0);
          f0 = __viaddmax_s16x2(f0, NEG1, hm0);
          const uint32_t t1 = (uint32_t)(hi1 * 65536 + (int32_t)(hd1 + (uint32_t)lo1));
          const uint32_t h1 = __vimax3_s16x2(t1, e0, f1);
          hd1 = h0;
          const uint32_t hm1 = __viaddmax_s16x2(h1, neg7, TWO2);
          E[c] = __viaddmax_s16x2(e0, NEG1, hm1);
          f1 = __viaddmax_s16x2(f1, NEG1, hm1);
          Hp[c] = h1;
          h0v[c] = h0;
          if (c & 1) { rb0 = __vimax3_s16x2(rb0, h0v[c - 1], h0); rb1 = __vimax3_s16x2(rb1, Hp[c - 1], h1); }
          else if (c == C - 1) { rb0 = __vimax3_s16x2(rb0, h0, h0); rb1 = __vimax3_s16x2(rb1, h1, h1); }
        };
        #pragma unroll
        for (int ch = 0; ch < L::CHF; ch++) {
          int32_t l0[4], l1[4], u0[4], u1[4];
          asm volatile("ld.shared.v4.s32 {%0, %1, %2, %3}, [%4];" : "=r"(l0[0]), "=r"(l0[1]), "=r"(l0[2]), "=r"(l0[3]) : "r"(r0l + (uint32_t)(ch * G * 16)));
          asm volatile("ld.shared.v4.s32 {%0, %1, %2, %3}, [%4];" : "=r"(u0[0]), "=r"(u0[1]), "=r"(u0[2]), "=r"(u0[3]) : "r"(r0h + (uint32_t)(ch * G * 16)));
          asm volatile("ld.shared.v4.s32 {%0, %1, %2, %3}, [%4];" : "=r"(l1[0]), "=r"(l1[1]), "=r"(l1[2]), "=r"(l1[3]) : "r"(r1l + (uint32_t)(ch * G * 16)));
          asm volatile("ld.shared.v4.s32 {%0, %1, %2, %3}, [%4];" : "=r"(u1[0]), "=r"(u1[1]), "=r"(u1[2]), "=r"(u1[3]) : "r"(r1h + (uint32_t)(ch * G * 16)));
          #pragma unroll
          for (int k = 0; k < 4; k++) cell(4 * ch + k, l0[k], u0[k], l1[k], u1[k]);
        }
        #pragma unroll
        for (int k = 0; k < L::TAIL; k++) {
          int32_t a0, b0, a1, b1;
          asm volatile("ld.shared.s32 %0, [%1];" : "=r"(a0) : "r"(tail_base + q0l * (uint32_t)L::TAIL_ROWB + (uint32_t)(k * 4)));
          asm volatile("ld.shared.s32 %0, [%1];" : "=r"(b0) : "r"(tail_base + q0h * (uint32_t)L::TAIL_ROWB + (uint32_t)(k * 4)));
          asm volatile("ld.shared.s32 %0, [%1];" : "=r"(a1) : "r"(tail_base + q1l * (uint32_t)L::TAIL_ROWB + (uint32_t)(k * 4)));
          asm volatile("ld.shared.s32 %0, [%1];" : "=r"(b1) : "r"(tail_base + q1h * (uint32_t)L::TAIL_ROWB + (uint32_t)(k * 4)));
          cell(4 * L::CHF + k, a0, b0, a1, b1);
        }
        lastH0 = h0v[C - 1]; lastH1 = Hp[C - 1];
        lastF0 = f0; lastF1 = f1;
        // first cell equal to the known maximum, per half (row 2*i2 before row 2*i2+1; within a row the lowest column)
        const uint32_t d0 = tgt - rb0;                      // per half >= 0 (S is the maximum), no borrow between the halves
        if ((d0 & 0xffffu) == 0u || d0 < 0x10000u) {
          if ((d0 & 0xffffu) == 0u) {
            int jj = 0;
            #pragma unroll
            for (int c = C - 1; c >= 0; c--) if ((h0v[c] & 0xffffu) == (tgt & 0xffffu)) jj = c;
            key_lo = ((uint32_t)(1023 - 2 * i2) << 10) | (uint32_t)(1023 - (g * C + jj));
            tgt |= 0x7FFFu;
          }
          if (d0 < 0x10000u) {
            int jj = 0;
            #pragma unroll
            for (int c = C - 1; c >= 0; c--) if ((h0v[c] >> 16) == (tgt >> 16)) jj = c;
            key_hi = ((uint32_t)(1023 - 2 * i2) << 10) | (uint32_t)(1023 - (g * C + jj));
            tgt |= 0x7FFF0000u;
          }
        }
        const uint32_t d1 = tgt - rb1;
        if ((d1 & 0xffffu) == 0u || d1 < 0x10000u) {
          if ((d1 & 0xffffu) == 0u) {
            int jj = 0;
            #pragma unroll
            for (int c = C - 1; c >= 0; c--) if ((Hp[c] & 0xffffu) == (tgt & 0xffffu)) jj = c;
            key_lo = ((uint32_t)(1023 - (2 * i2 + 1)) << 10) | (uint32_t)(1023 - (g * C + jj));
            tgt |= 0x7FFFu;
          }
          if (d1 < 0x10000u) {
            int jj = 0;
            #pragma unroll
            for (int c = C - 1; c >= 0; c--) if ((Hp[c] >> 16) == (tgt >> 16)) jj = c;
            key_hi = ((uint32_t)(1023 - (2 * i2 + 1)) << 10) | (uint32_t)(1023 - (g * C + jj));
            tgt |= 0x7FFF0000u;
          }
        }
      }
    }
    #pragma unroll
    for (int d = G / 2; d; d >>= 1) {
      key_lo = max(key_lo, __shfl_xor_sync(FULL, key_lo, d, G));
      key_hi = max(key_hi, __shfl_xor_sync(FULL, key_hi, d, G));
    }
    if (live && g == 0) {
      out[ka] = ((uint32_t)wtasks[ka].b << 20) | key_lo;
      if (kb != NO_WINNER) out[kb] = ((uint32_t)wtasks[kb].b << 20) | key_hi;
    }
    __syncwarp();
  }
}

template <int G, int C>
int launch_sw_track2(const RefDev& ref, const ReadsDev& reads, AlignBuffers& ab, int sm_count, cudaStream_t st) {
  constexpr int GPB = (32 / G) * (SW_THREADS / 32);
  const int ml = qarr_stride(reads.max_len);                     // qarr bytes per task: one u16 per row pair
  const size_t smem = (size_t)GPB * (ProfLayout<G, C>::SLOTB + ml);
  auto k = sw_track2_kernel<G, C>;
  static std::atomic<unsigned long long> optin{0};
  int per_sm = 0;
  { const int rc = smem_optin(k, 200 * 1024, optin); if (rc) return rc; }
  CATT_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k, SW_THREADS, smem));
  if (per_sm < 1) per_sm = 1;
  const int64_t want = ((reads.n + ref.nrec) / 2 + GPB) / GPB;
  const int grid = (int)std::max<int64_t>(1, std::min<int64_t>((int64_t)sm_count * per_sm, want));
  k<<<grid, SW_THREADS, smem, st>>>(ref, reads, ab.wtasks, ab.worder, ab.whist + 2 * MAX_REC, ab.wres, ml, ab.counters + CTR_WORK_TRACK);
  CATT_CUDA(cudaGetLastError());
  return CATT_OK;
}

// end cells of the winners in ab.wtasks (scores in SwTask::b): grouping by record + the profile TRACK kernel
int dispatch_track2(const RefDev& ref, const ReadsDev& reads, AlignBuffers& ab, int sm_count, cudaStream_t st, int* launches) {
  const int64_t n = reads.n;
  uint32_t* hist = ab.whist;                       // [MAX_REC] counts, [MAX_REC] cursors, [1] slots
  CATT_CUDA(cudaMemsetAsync(hist, 0, MAX_REC * sizeof(uint32_t), st));
  whist_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(n, ab.wtasks, ab.counters, hist);
  wscan_kernel<<<1, MAX_REC, 0, st>>>(ref.nrec, hist, hist + MAX_REC, ab.worder, hist + 2 * MAX_REC);
  wscatter_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(n, ab.wtasks, ab.counters, hist + MAX_REC, ab.worder);
  CATT_CUDA(cudaGetLastError());
  *launches += 4;
  const int len = ref.maxlen;
  if (len <= 32) return launch_sw_track2<4, 8>(ref, reads, ab, sm_count, st);
  if (len <= 48) return launch_sw_track2<4, 12>(ref, reads, ab, sm_count, st);
  if (len <= 64) return launch_sw_track2<8, 8>(ref, reads, ab, sm_count, st);
  if (len <= 100) return launch_sw_track2<4, 25>(ref, reads, ab, sm_count, st);
  if (len <= 104) return launch_sw_track2<8, 13>(ref, reads, ab, sm_count, st);
  if (len <= 160) return launch_sw_track2<8, 20>(ref, reads, ab, sm_count, st);
  if (len <= 320) return launch_sw_track2<16, 20>(ref, reads, ab, sm_count, st);
  set_error("record length %d exceeds the SW kernel's tiling", len);
  return CATT_E_LIMIT;
}

template <bool TRACK>
int dispatch_sw(const RefDev& ref, const ReadsDev& reads, const SwTask* tasks, const uint32_t* ntasks_dev, int64_t cap, uint32_t* res,
                unsigned long long* work, int sm_count, cudaStream_t st) {
  const int n = ref.maxlen;
  const char* t32 = n <= 32 ? getenv("CATT_TILE32") : nullptr;              // A/B switch, see INTEGRATION.md
  const int tile32 = t32 ? atoi(t32) : 4;
  if (n <= 32 && tile32 == 2) return launch_sw_t<2, 16, TRACK>(ref, reads, tasks, ntasks_dev, cap, res, work, sm_count, st);
  if (n <= 32 && tile32 == 1) return launch_sw_t<1, 32, TRACK>(ref, reads, tasks, ntasks_dev, cap, res, work, sm_count, st);
  if (n <= 32) return launch_sw_t<4, 8, TRACK>(ref, reads, tasks, ntasks_dev, cap, res, work, sm_count, st);
  if (n <= 48) return launch_sw_t<4, 12, TRACK>(ref, reads, tasks, ntasks_dev, cap, res, work, sm_count, st);
  if (n <= 64) return launch_sw_t<8, 8, TRACK>(ref, reads, tasks, ntasks_dev, cap, res, work, sm_count, st);
  if (n <= 100) return launch_sw_t<4, 25, TRACK>(ref, reads, tasks, ntasks_dev, cap, res, work, sm_count, st);
  if (n <= 104) return launch_sw_t<8, 13, TRACK>(ref, reads, tasks, ntasks_dev, cap, res, work, sm_count, st);
  if (n <= 160) return launch_sw_t<8, 20, TRACK>(ref, reads, tasks, ntasks_dev, cap, res, work, sm_count, st);
  if (n <= 320) return launch_sw_t<16, 20, TRACK>(ref, reads, tasks, ntasks_dev, cap, res, work, sm_count, st);
  set_error("record length %d exceeds the SW kernel's tiling", n);
  return CATT_E_LIMIT;
}

// ---------------------------------------------------------------------------------------------------------------
// selection.  One thread per read: best score over the candidates, bwa's tie-break, the winner's task for the
// TRACK pass.
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ int sub_score(uint32_t a, uint32_t b) { return (a > 3 || b > 3) ? -1 : (a == b ? 1 : -2); }

// MODE 0: the whole selection.  MODE 1 / 2 split it where bwa's tie-break enters (hash_64(read ordinal + rank among the tied)):
// MODE 1 scores the candidates and leaves, per read, the best score, the candidate count and the bitmap of the candidates that
// reach it; MODE 2 picks the winner from those once the read's global ordinal is known.  A GPU group that ingests a file learns
// the ordinals only when every rank has counted its reads, and runs seeding + SW + MODE 1 while the file is still being read.
template <int MODE>
__global__ void select_kernel(RefDev ref, ReadsDev reads, const uint32_t* __restrict__ candbits, const uint32_t* __restrict__ canonbits,
                              const SwTask* __restrict__ tasks,
                              const uint32_t* __restrict__ task_off, const uint32_t* __restrict__ task_score, int64_t cap_tasks,
                              int64_t first_ordinal, int min_score, catt_aln* __restrict__ recs, SwTask* __restrict__ wtasks,
                              unsigned long long* __restrict__ counters, uint32_t* __restrict__ tied, int16_t* __restrict__ mxs,
                              int16_t* __restrict__ ncands, uint32_t* __restrict__ wres) {
  const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  unsigned long long my_cand = 0, my_cells = 0, my_hit = 0, my_comp = 0;
  if (MODE == 2) {
    if (r < reads.n) {
      catt_aln a;
      a.ordinal = (int32_t)(first_ordinal + r);
      a.rid = -1; a.strand = 0; a.mapped = 0; a.qb = 0; a.m_end = 0; a.rb = 0; a.qe = 0; a.re = 0; a.nm = 0; a.mi = 0; a.ntied = 0;
      const int mx = mxs[r], ncand = ncands[r];
      a.as = (int16_t)mx; a.ncand = (int16_t)ncand;
      if (ncand > 0 && mx >= min_score) {
        const uint64_t ord = (uint64_t)(first_ordinal + r);
        int ntied = 0, win_c = -1;
        uint64_t win_h = ~0ull;
        auto consider = [&](int c) {
          const uint64_t h = hash_64(ord + (uint64_t)ntied);
          ntied++;
          if (h < win_h) { win_h = h; win_c = c; }
        };
        for (int w = 0; w < ref.CW; w++) {                    // T order: forward records ascending, then reverse-strand records descending
          uint32_t b = tied[r * ref.CW + w] & 0x55555555u;
          while (b) { const int bit = __ffs(b) - 1; b &= b - 1; consider(w * 32 + bit); }
        }
        for (int w = ref.CW - 1; w >= 0; w--) {
          uint32_t b = tied[r * ref.CW + w] & 0xAAAAAAAAu;
          while (b) { const int bit = 31 - __clz(b); b &= ~(1u << bit); consider(w * 32 + bit); }
        }
        a.rid = win_c >> 1; a.strand = (int16_t)(win_c & 1); a.mapped = 1; a.ntied = (int16_t)ntied;
        my_hit = 1;
        const unsigned long long k = atomicAdd(&counters[7], 1ull);
        wtasks[k] = SwTask{(uint32_t)r, (uint16_t)win_c, (uint16_t)mx};
      }
      recs[r] = a;
    }
  } else if (r < reads.n && (int64_t)task_off[reads.n] <= cap_tasks) {
    const uint32_t t0 = task_off[r], t1 = task_off[r + 1];
    const int m = reads.len[r];
    catt_aln a;
    a.ordinal = (int32_t)(first_ordinal + r);
    a.rid = -1; a.strand = 0; a.mapped = 0; a.as = 0; a.qb = 0; a.m_end = 0; a.rb = 0; a.qe = 0; a.re = 0; a.nm = 0; a.mi = 0;
    a.ncand = 0; a.ntied = 0;
    int mx = 0;
    for (uint32_t t = t0; t < t1; t++) {
      const SwTask tk = tasks[t];
      const uint32_t sc = task_score[t];
      my_comp += (unsigned long long)m * ref.rec_len[tk.a >> 1];
      if (ref.direct) { mx = max(mx, (int)(sc >> 20)); continue; }          // one candidate per task: its TRACK key (pack_key)
      mx = max(mx, (int)(sc & 0xffff));
      if (tk.b != 0xFFFF) { mx = max(mx, (int)(sc >> 16)); my_comp += (unsigned long long)m * ref.rec_len[tk.b >> 1]; }
    }
    // candidates in T order (forward records ascending, then reverse-strand records descending): rank among the tied
    const uint64_t ord = (uint64_t)(first_ordinal + r);
    int ncand = 0, ntied = 0, win_c = -1;
    uint64_t win_h = ~0ull;
    uint32_t win_key = 0;
    // expand_kernel emitted the tasks in bit order of the canonical bitmap, forward strand first, two candidates per task: the
    // score of canonical candidate cc sits in task t0 (+ forward tasks) + rank/2, half rank%2, rank = its position among the
    // canonical candidates of its strand - a popcount, not a search
    uint16_t pre[2][MAX_CW];
    uint32_t nfw = 0, nrv = 0;
    for (int w = 0; w < ref.CW; w++) {
      const uint32_t b = canonbits[r * ref.CW + w];
      pre[0][w] = (uint16_t)nfw; pre[1][w] = (uint16_t)nrv;
      nfw += __popc(b & 0x55555555u); nrv += __popc(b & 0xAAAAAAAAu);
      if (MODE == 1) tied[r * ref.CW + w] = 0;
    }
    const uint32_t rv_base = ref.direct ? t0 + nfw : t0 + (nfw + 1) / 2;
    auto consider = [&](int c) {
      ncand++;
      my_cells += (unsigned long long)m * ref.rec_len[c >> 1];
      const int cc = ref.canon[c];
      const int st = cc & 1, w = cc >> 5;
      const uint32_t below = canonbits[r * ref.CW + w] & (st ? 0xAAAAAAAAu : 0x55555555u) & ((1u << (cc & 31)) - 1u);
      const uint32_t rank = pre[st][w] + __popc(below);
      const uint32_t ts = task_score[(st ? rv_base : t0) + (ref.direct ? rank : rank >> 1)];
      const int sc = ref.direct ? (int)(ts >> 20) : (int)((rank & 1) ? ts >> 16 : ts & 0xffff);
      if (sc == mx) {
        if (MODE == 1) {
          tied[r * ref.CW + (c >> 5)] |= 1u << (c & 31);
        } else {
          const uint64_t h = hash_64(ord + (uint64_t)ntied);
          ntied++;
          if (h < win_h) { win_h = h; win_c = c; win_key = ts; }
        }
      }
    };
    if (t1 > t0) {
      for (int w = 0; w < ref.CW; w++) {
        uint32_t b = candbits[r * ref.CW + w] & 0x55555555u;
        while (b) { const int bit = __ffs(b) - 1; b &= b - 1; consider(w * 32 + bit); }
      }
      for (int w = ref.CW - 1; w >= 0; w--) {
        uint32_t b = candbits[r * ref.CW + w] & 0xAAAAAAAAu;
        while (b) { const int bit = 31 - __clz(b); b &= ~(1u << bit); consider(w * 32 + bit); }
      }
    }
    my_cand = ncand;
    if (MODE == 1) {
      mxs[r] = (int16_t)mx; ncands[r] = (int16_t)ncand;
    } else {
      a.ncand = (int16_t)ncand;
      a.as = (int16_t)mx;
      if (ncand > 0 && mx >= min_score) {
        a.rid = win_c >> 1; a.strand = (int16_t)(win_c & 1); a.mapped = 1; a.ntied = (int16_t)ntied;
        my_hit = 1;
        const unsigned long long k = atomicAdd(&counters[7], 1ull);
        wtasks[k] = SwTask{(uint32_t)r, (uint16_t)win_c, (uint16_t)mx};      // b = the winner's score (dispatch_track2)
        if (ref.direct) wres[k] = win_key;           // the winner's end cell is already known: no TRACK pass (launch_select)
      }
      recs[r] = a;
    }
  }
  // block-level reduction of the work counters
  __shared__ unsigned long long sh[4];
  if (threadIdx.x < 4) sh[threadIdx.x] = 0;
  __syncthreads();
  #pragma unroll
  for (int d = 16; d; d >>= 1) {
    my_cand += __shfl_xor_sync(FULL, my_cand, d);
    my_cells += __shfl_xor_sync(FULL, my_cells, d);
    my_hit += __shfl_xor_sync(FULL, my_hit, d);
    my_comp += __shfl_xor_sync(FULL, my_comp, d);
  }
  if ((threadIdx.x & 31) == 0) { atomicAdd(&sh[0], my_cand); atomicAdd(&sh[1], my_cells); atomicAdd(&sh[2], my_hit); atomicAdd(&sh[3], my_comp); }
  __syncthreads();
  if (threadIdx.x == 0) {
    atomicAdd(&counters[0], sh[0]); atomicAdd(&counters[1], sh[1]); atomicAdd(&counters[3], sh[2]); atomicAdd(&counters[8], sh[3]);
  }
}

// CIGAR facts of the winner from its end cell (ungapped fast path): walk back along the diagonal until the suffix sums
// to AS; otherwise the read goes to the traceback kernel.  One thread per winner.
__global__ void finish_kernel(RefDev ref, ReadsDev reads, const SwTask* __restrict__ wtasks, const uint32_t* __restrict__ wres,
                              catt_aln* __restrict__ recs, uint32_t* __restrict__ gapped_list, unsigned long long* __restrict__ counters) {
  const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= (int64_t)counters[7]) return;
  const SwTask tk = wtasks[k];
  const uint32_t r = tk.read;
  const uint32_t key = wres[k];
  catt_aln a = recs[r];
  const int mx = a.as, strand = a.strand, m = reads.len[r];
  const int ei = 1023 - (int)((key >> 10) & 1023), ej = 1023 - (int)(key & 1023);
  a.qe = (int16_t)(ei + 1); a.re = (int16_t)(ej + 1);
  const uint32_t* __restrict__ w = reads.words + reads.off[r];
  const uint32_t* __restrict__ nm = w + read_base_words(m);
  const uint8_t* __restrict__ tb = ref.Tb + ref.rec_off[a.rid];
  int sum = 0, mism = 0, d = 0;
  bool ok = false;
  const int lim = min(ei, ej);
  for (; d <= lim; d++) {
    const uint32_t q = read_code(w, nm, m, ei - d, strand);
    const uint32_t tt = tb[ej - d];
    const int s = sub_score(q, tt);
    sum += s; mism += (s != 1);
    if (sum == mx) { ok = true; break; }
  }
  if (ok) {
    a.qb = (int16_t)(ei - d); a.rb = (int16_t)(ej - d);
    a.m_end = (int16_t)(ei + 1);
    a.nm = (int16_t)mism; a.mi = (int16_t)(d + 1);
  } else {
    const unsigned long long g = atomicAdd(&counters[2], 1ull);
    gapped_list[g] = r;
  }
  recs[r] = a;
}

// ---------------------------------------------------------------------------------------------------------------
// traceback for the rare gapped winners: one warp per read.  The sub-matrix up to the end cell is filled by
// anti-diagonals (three rolling H diagonals, two E/F) and only a 4-bit direction code per cell is kept in shared
// memory; then one lane walks back.  Per-cell preference M, then D (record base skipped), then I; a gap prefers
// "open" over "extend" on a tie; stops at the first H == 0.
// ---------------------------------------------------------------------------------------------------------------
constexpr int TB_NEG = -30000;
__global__ void __launch_bounds__(32) traceback_kernel(RefDev ref, ReadsDev reads, const uint32_t* __restrict__ gapped_list,
                                                       const unsigned long long* __restrict__ counters, catt_aln* __restrict__ recs) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int lane = threadIdx.x;
  int16_t* Hb = reinterpret_cast<int16_t*>(smem_raw);          // 3 x (MAX_READ+2)
  int16_t* Eb = Hb + 3 * (MAX_READ + 2);                       // 2 x
  int16_t* Fb = Eb + 2 * (MAX_READ + 2);                       // 2 x
  uint8_t* q = reinterpret_cast<uint8_t*>(Fb + 2 * (MAX_READ + 2));
  uint8_t* dir = q + MAX_READ;
  const int64_t n_gapped = (int64_t)counters[2];
  for (int64_t it = blockIdx.x; it < n_gapped; it += gridDim.x) {
    const uint32_t r = gapped_list[it];
    catt_aln a = recs[r];
    const int m = reads.len[r];
    const int rows = a.qe, cols = a.re;                 // sub-matrix [0,rows) x [0,cols); end cell = (rows-1, cols-1)
    const uint32_t* __restrict__ w = reads.words + reads.off[r];
    const uint32_t* __restrict__ nm = w + read_base_words(m);
    const uint8_t* __restrict__ tb = ref.Tb + ref.rec_off[a.rid];
    for (int i = lane; i < rows; i += 32) q[i] = (uint8_t)read_code(w, nm, m, i, a.strand);
    __syncwarp();
    for (int d = 0; d <= rows + cols; d++) {            // cells (i,j), 1-based with a 0 border, i + j == d
      int16_t* Hc = Hb + (d % 3) * (MAX_READ + 2);
      const int16_t* H1 = Hb + ((d + 2) % 3) * (MAX_READ + 2);
      const int16_t* H2 = Hb + ((d + 1) % 3) * (MAX_READ + 2);
      int16_t* Ec = Eb + (d & 1) * (MAX_READ + 2);
      const int16_t* E1 = Eb + ((d + 1) & 1) * (MAX_READ + 2);
      int16_t* Fc = Fb + (d & 1) * (MAX_READ + 2);
      const int16_t* F1 = Fb + ((d + 1) & 1) * (MAX_READ + 2);
      const int ilo = max(0, d - cols), ihi = min(rows, d);
      for (int i = ilo + lane; i <= ihi; i += 32) {
        const int j = d - i;
        if (i == 0 || j == 0) { Hc[i] = 0; Ec[i] = TB_NEG; Fc[i] = TB_NEG; continue; }
        const int eo = H1[i - 1] - 7, ee = E1[i - 1] - 1;      // vertical: consumes a read base (I)
        const int fo = H1[i] - 7, fe = F1[i] - 1;              // horizontal: consumes a record base (D)
        const int e = max(eo, ee), f = max(fo, fe);
        const int dg = H2[i - 1] + sub_score(q[i - 1], tb[j - 1]);
        const int h = max(max(0, dg), max(e, f));
        const int src = h == 0 ? 0 : (h == dg ? 1 : (h == f ? 2 : 3));
        dir[(size_t)(i - 1) * cols + (j - 1)] = (uint8_t)(src | ((fe > fo) << 2) | ((ee > eo) << 3));
        Hc[i] = (int16_t)h; Ec[i] = (int16_t)max(e, TB_NEG); Fc[i] = (int16_t)max(f, TB_NEG);
      }
      __syncwarp();
    }
    if (lane == 0) {
      int i = rows, j = cols, state = 0, nmm = 0, mi = 0, run_m = 0;
      // ops are visited end -> start; the first M block is the LAST run of M's visited
      while (i > 0 && j > 0) {
        const uint8_t dc = dir[(size_t)(i - 1) * cols + (j - 1)];
        if (state == 0) {
          const int src = dc & 3;
          if (src == 0) break;
          if (src == 1) { nmm += (sub_score(q[i - 1], tb[j - 1]) != 1); mi++; run_m++; i--; j--; }
          else if (src == 2) { state = 1; run_m = 0; }
          else { state = 2; run_m = 0; }
        } else if (state == 1) {                        // D
          nmm++;
          if (!(dc & 4)) state = 0;
          j--;
        } else {                                        // I
          nmm++; mi++;
          if (!(dc & 8)) state = 0;
          i--;
        }
      }
      a.qb = (int16_t)i; a.rb = (int16_t)j; a.m_end = (int16_t)(i + run_m); a.nm = (int16_t)nmm; a.mi = (int16_t)mi;
      recs[r] = a;
    }
    __syncwarp();
  }
}

// ---------------------------------------------------------------------------------------------------------------
// integer-ALU peak microbenchmark: a dependent-free stream of the packed DPX add-max the DP is made of
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) alu_peak_kernel(uint32_t* out, int iters) {
  uint32_t a[8];
  #pragma unroll
  for (int k = 0; k < 8; k++) a[k] = threadIdx.x * 2654435761u + k;
  const uint32_t b = 0x00010001u + (blockIdx.x & 1), c = 0x00030002u;
  for (int it = 0; it < iters; it++) {
    #pragma unroll
    for (int k = 0; k < 8; k++) a[k] = __viaddmax_s16x2(a[k], b, c);
    #pragma unroll
    for (int k = 0; k < 8; k++) a[k] = __viaddmax_s16x2_relu(a[k], c, a[(k + 1) & 7]);
  }
  uint32_t s = 0;
  #pragma unroll
  for (int k = 0; k < 8; k++) s ^= a[k];
  if (s == 0x12345678u) out[0] = s;
}

// SW of explicit (read i, record rid[i], strand[i]) pairs — parity hook around the same kernel
__global__ void make_pair_tasks(int64_t n, const int32_t* rid, const int32_t* strand, SwTask* tasks) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) tasks[i] = SwTask{(uint32_t)i, (uint16_t)(rid[i] * 2 + (strand[i] & 1)), (uint16_t)0xFFFF};
}
__global__ void pair_scores_kernel(int64_t n, const uint32_t* res, SwTask* tasks) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) tasks[i].b = (uint16_t)(res[i] >> 20);
}
__global__ void unpack_pair_res(int64_t n, const uint32_t* res, int32_t* out3) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) {
    const uint32_t k = res[i];
    out3[3 * i] = (int32_t)(k >> 20);
    out3[3 * i + 1] = 1023 - (int32_t)((k >> 10) & 1023);
    out3[3 * i + 2] = 1023 - (int32_t)(k & 1023);
  }
}

}  // namespace

int ensure_task_buffers(AlignBuffers& ab, int64_t want) {
  if (want <= ab.cap_tasks) return CATT_OK;
  if (ab.tasks) cudaFree(ab.tasks);
  if (ab.task_score) cudaFree(ab.task_score);
  ab.tasks = nullptr; ab.task_score = nullptr; ab.cap_tasks = 0;
  CATT_CUDA(cudaMalloc(&ab.tasks, want * sizeof(SwTask)));
  CATT_CUDA(cudaMalloc(&ab.task_score, want * sizeof(uint32_t)));
  ab.cap_tasks = want;
  return CATT_OK;
}

// scan of the per-read task counts + task expansion; no host round trip (the task count stays on the device)
int launch_expand(const RefDev& ref, const ReadsDev& reads, AlignBuffers& ab, cudaStream_t st, int* launches) {
  const int64_t n = reads.n;
  int rc;
  if ((rc = ensure_task_buffers(ab, n * 10 + 4096))) return rc;     // grown (and the stage re-run) on overflow
  size_t need = 0;
  CATT_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, need, ab.ntask, ab.task_off, (int)(n + 1), st));
  if (need > ab.cub_tmp_bytes) {
    if (ab.cub_tmp) cudaFree(ab.cub_tmp);
    CATT_CUDA(cudaMalloc(&ab.cub_tmp, need));
    ab.cub_tmp_bytes = need;
  }
  CATT_CUDA(cub::DeviceScan::ExclusiveSum(ab.cub_tmp, need, ab.ntask, ab.task_off, (int)(n + 1), st));
  expand_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(ref, n, ab.canonbits, ab.task_off, ab.tasks, ab.cap_tasks, ab.counters);
  CATT_CUDA(cudaGetLastError());
  *launches += 3;      // cub scan = 2 kernels
  return CATT_OK;
}

int launch_sw(const RefDev& ref, const ReadsDev& reads, AlignBuffers& ab, int sm_count, cudaStream_t st, int* launches) {
  *launches += 1;
  // a reference set whose reads have one or two candidates each (the J sets) skips the bulk pass: its packed halves would be half
  // empty and the winner would be scored a second time; every candidate goes through the TRACK kernel once, two reads per task
  if (ref.direct) return dispatch_sw<true>(ref, reads, ab.tasks, ab.task_off + reads.n, ab.cap_tasks, ab.task_score, ab.counters + CTR_WORK_SW, sm_count, st);
  return dispatch_sw_bulk(ref, reads, ab.tasks, ab.task_off + reads.n, ab.cap_tasks, ab.task_score, ab.counters + CTR_WORK_SW, sm_count, st);
}

// selection, the TRACK pass over the winners, CIGAR facts, traceback of the rare gapped winners
// TRACK pass on the winners, CIGAR facts, traceback of the gapped ones: everything after the winner of every read is known
static int launch_after_select(const RefDev& ref, const ReadsDev& reads, AlignBuffers& ab, catt_aln* recs, bool track, int sm_count, cudaStream_t st,
                               int* launches) {
  const bool old_track = getenv("CATT_OLD_TRACK") != nullptr;              // A/B switch: the PRMT kernel on unsorted winners
  int rc = CATT_OK;
  if (track && old_track)
    rc = dispatch_sw<true>(ref, reads, ab.wtasks, reinterpret_cast<const uint32_t*>(ab.counters + 7), reads.n, ab.wres, ab.counters + CTR_WORK_TRACK, sm_count, st);
  else if (track)
    rc = dispatch_track2(ref, reads, ab, sm_count, st, launches);
  if (rc) return rc;
  finish_kernel<<<(unsigned)((reads.n + 127) / 128), 128, 0, st>>>(ref, reads, ab.wtasks, ab.wres, recs, ab.gapped_list, ab.counters);
  CATT_CUDA(cudaGetLastError());
  const size_t smem = (size_t)7 * (MAX_READ + 2) * sizeof(int16_t) + MAX_READ + (size_t)reads.max_len * ref.maxlen + 16;
  if (smem > 220 * 1024) {
    set_error("traceback tile (%zu bytes) exceeds shared memory", smem);
    return CATT_E_LIMIT;
  }
  static std::atomic<unsigned long long> optin{0};
  if ((rc = smem_optin(traceback_kernel, 220 * 1024, optin))) return rc;
  traceback_kernel<<<sm_count * 4, 32, smem, st>>>(ref, reads, ab.gapped_list, ab.counters, recs);
  CATT_CUDA(cudaGetLastError());
  *launches += track ? 3 : 2;
  return CATT_OK;
}

int launch_select(const RefDev& ref, const ReadsDev& reads, AlignBuffers& ab, catt_aln* recs, int64_t first_ordinal, int min_score,
                  int sm_count, cudaStream_t st, int* launches) {
  if (reads.n == 0) return CATT_OK;
  select_kernel<0><<<(unsigned)((reads.n + 127) / 128), 128, 0, st>>>(ref, reads, ab.candbits, ab.canonbits, ab.tasks, ab.task_off, ab.task_score,
                                                                      ab.cap_tasks, first_ordinal, min_score, recs, ab.wtasks, ab.counters, nullptr, nullptr, nullptr, ab.wres);
  CATT_CUDA(cudaGetLastError());
  *launches += 1;
  return launch_after_select(ref, reads, ab, recs, !ref.direct, sm_count, st, launches);
}

// the selection split at the tie-break (select_kernel MODE 1 / 2): `tied` = reads.n x CW words, `mxs` / `ncands` = reads.n each
int launch_select_pre(const RefDev& ref, const ReadsDev& reads, AlignBuffers& ab, uint32_t* tied, int16_t* mxs, int16_t* ncands, cudaStream_t st, int* launches) {
  if (reads.n == 0) return CATT_OK;
  select_kernel<1><<<(unsigned)((reads.n + 127) / 128), 128, 0, st>>>(ref, reads, ab.candbits, ab.canonbits, ab.tasks, ab.task_off, ab.task_score,
                                                                      ab.cap_tasks, 0, 0, nullptr, nullptr, ab.counters, tied, mxs, ncands, nullptr);
  CATT_CUDA(cudaGetLastError());
  *launches += 1;
  return CATT_OK;
}
int launch_select_post(const RefDev& ref, const ReadsDev& reads, AlignBuffers& ab, uint32_t* tied, int16_t* mxs, int16_t* ncands, catt_aln* recs,
                       int64_t first_ordinal, int min_score, int sm_count, cudaStream_t st, int* launches) {
  if (reads.n == 0) return CATT_OK;
  select_kernel<2><<<(unsigned)((reads.n + 127) / 128), 128, 0, st>>>(ref, reads, nullptr, nullptr, nullptr, nullptr, nullptr, 0, first_ordinal, min_score,
                                                                      recs, ab.wtasks, ab.counters, tied, mxs, ncands, nullptr);
  CATT_CUDA(cudaGetLastError());
  *launches += 1;
  return launch_after_select(ref, reads, ab, recs, true, sm_count, st, launches);     // the front half's keys are gone: TRACK pass on the winners
}

int launch_sw_pairs(const RefDev& ref, const ReadsDev& reads, const int32_t* d_rid, const int32_t* d_strand, int32_t* d_out3,
                    int sm_count, cudaStream_t st) {
  const int64_t n = reads.n;
  if (n == 0) return CATT_OK;
  SwTask* tasks = nullptr;
  uint32_t* res = nullptr;
  uint32_t* d_n = nullptr;                       // [0] task count, [2..3] the kernel's work cursor
  CATT_CUDA(cudaMalloc(&tasks, n * sizeof(SwTask)));
  CATT_CUDA(cudaMalloc(&res, n * 2 * sizeof(uint32_t)));
  CATT_CUDA(cudaMalloc(&d_n, 4 * sizeof(uint32_t)));
  const uint32_t n32[4] = {(uint32_t)n, 0, 0, 0};
  CATT_CUDA(cudaMemcpyAsync(d_n, n32, sizeof n32, cudaMemcpyHostToDevice, st));
  make_pair_tasks<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(n, d_rid, d_strand, tasks);
  int rc = dispatch_sw<true>(ref, reads, tasks, d_n, n, res, reinterpret_cast<unsigned long long*>(d_n + 2), sm_count, st);
  // the scores are known now: the same pairs through the grouped profile kernel (what Stage A runs on the V winners); its end
  // cells are the ones returned
  AlignBuffers ab2;
  if (rc == CATT_OK && !getenv("CATT_OLD_TRACK")) {
    unsigned long long hc[N_COUNTERS] = {0};
    hc[7] = (unsigned long long)n;
    int launches = 0;
    if (cudaMalloc(&ab2.counters, sizeof hc) != cudaSuccess || cudaMalloc(&ab2.worder, (n + MAX_REC + 2) * sizeof(uint32_t)) != cudaSuccess ||
        cudaMalloc(&ab2.whist, (2 * MAX_REC + 4) * sizeof(uint32_t)) != cudaSuccess || cudaMalloc(&ab2.wres, n * sizeof(uint32_t)) != cudaSuccess) {
      set_error("sw_pairs: out of device memory"); rc = CATT_E_CUDA;
    } else {
      cudaMemcpyAsync(ab2.counters, hc, sizeof hc, cudaMemcpyHostToDevice, st);
      pair_scores_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(n, res, tasks);
      ab2.wtasks = tasks;
      rc = dispatch_track2(ref, reads, ab2, sm_count, st, &launches);
      if (rc == CATT_OK) cudaMemcpyAsync(res, ab2.wres, n * sizeof(uint32_t), cudaMemcpyDeviceToDevice, st);
    }
  }
  if (rc == CATT_OK) {
    unpack_pair_res<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(n, res, d_out3);
    cudaError_t e = cudaStreamSynchronize(st);
    if (e != cudaSuccess) rc = cuda_fail(e, "sw_pairs", __FILE__, __LINE__);
  }
  cudaFree(tasks); cudaFree(res); cudaFree(d_n);
  cudaFree(ab2.counters); cudaFree(ab2.worder); cudaFree(ab2.whist); cudaFree(ab2.wres);
  return rc;
}

int alu_peak(int sm_count, cudaStream_t st, double* ops_per_s, double* ms_out) {
  uint32_t* d = nullptr;
  CATT_CUDA(cudaMalloc(&d, 16));
  const int iters = 4096, grid = sm_count * 8, block = 256;
  cudaEvent_t e0, e1;
  CATT_CUDA(cudaEventCreate(&e0)); CATT_CUDA(cudaEventCreate(&e1));
  alu_peak_kernel<<<grid, block, 0, st>>>(d, 64);
  float best = 1e30f;
  for (int rep = 0; rep < 5; rep++) {
    CATT_CUDA(cudaEventRecord(e0, st));
    alu_peak_kernel<<<grid, block, 0, st>>>(d, iters);
    CATT_CUDA(cudaEventRecord(e1, st));
    CATT_CUDA(cudaEventSynchronize(e1));
    float ms = 0;
    CATT_CUDA(cudaEventElapsedTime(&ms, e0, e1));
    best = std::min(best, ms);
  }
  const double ops = (double)grid * block * (double)iters * 16.0;   // packed s16x2 instructions (2 lanes each)
  *ops_per_s = ops / (best * 1e-3);
  *ms_out = best;
  cudaEventDestroy(e0); cudaEventDestroy(e1); cudaFree(d);
  return CATT_OK;
}

}  // namespace catt
