// finder.cuh — warp-cooperative finder! : 3-frame translation + Cys / Phe-Trp motif scan -> CDR3 nucleotide range.
//
// Replaces finder! (Jtool.jl:355-389) and both Extract_Motif2 methods (Jtool.jl:234-316):
//   for shift in 1:3  aa = translate(seq[shift : lgt-(lgt-shift+1)%3])
//     Cx = offsets of cmotif (+coffset), Fx = offsets of fmotif (+foffset)      [PCRE eachmatch: leftmost, non-overlapping]
//     exactly one empty -> fill it from the inner motif, code 2 (single-read version: the hard-coded lists of
//     Jtool.jl:288,294; array version: refg innerC (+0) / innerF (+2))
//     for xc descending, xf ascending: 6 <= xf-xc <= 34, not (7 <= xf - nextC <= 35), no '*' in aa[xc:xf] -> code 3,
//     range (xc-1)*3+1 : xf*3, break the inner loop only (so the leftmost qualifying C wins)
//   able |= code > 0; the first frame with code 3 sets cdr3.
// Motifs are compiled on the host into per-position tables of alternative sets (MotifDev).
#pragma once
#include "internal.cuh"

namespace catt {

constexpr int FINDER_MAX_AA = (MAX_READ + 2) / 3;          // 171
constexpr int FINDER_WORDS = (FINDER_MAX_AA + 31) / 32;    // 6

struct FinderScratch {        // per warp, shared memory
  uint8_t aa[FINDER_MAX_AA + 5];
  uint32_t cw[FINDER_WORDS], fw[FINDER_WORDS];
  uint8_t Cx[FINDER_MAX_AA], Fx[FINDER_MAX_AA];
};

// residue index: letter - 'A' (0..25), '*' = 26; table indexed b1*16+b2*4+b3 with A0 C1 G2 T3 (standard code)
static __device__ __constant__ char FINDER_CODON[65] = "KNKNTTTTRSRSIIMIQHQHPPPPRRRRLLLLEDEDAAAAGGGGVVVV*Y*YSSSS*CWCLFLF";

__device__ __forceinline__ bool motif_at(const MotifDev& m, const uint8_t* aa, int p) {
  unsigned long long live = ~0ull;                       // alternatives still matching
  for (int k = 0; k < m.len; k++) live &= m.alt[k][aa[p + k]];
  return live != 0;
}

// match bitmask over amino-acid positions (bit p = the motif matches at p), written to `words`; returns any-match
__device__ __forceinline__ bool motif_scan(const MotifDev& m, const uint8_t* aa, int naa, uint32_t* words, int lane) {
  bool any = false;
  #pragma unroll
  for (int c = 0; c < FINDER_WORDS; c++) {
    uint32_t b = 0;
    if (c * 32 < naa) {
      const int p = c * 32 + lane;
      const bool hit = m.nalt > 0 && p + m.len <= naa && motif_at(m, aa, p);
      b = __ballot_sync(0xffffffffu, hit);
    }
    if (lane == 0) words[c] = b;
    any |= (b != 0);
  }
  __syncwarp();
  return any;
}

__device__ __forceinline__ int next_bit(const uint32_t* w, int from) {
  int wi = from >> 5;
  if (wi >= FINDER_WORDS) return -1;
  uint32_t b = w[wi] & (0xffffffffu << (from & 31));
  while (true) {
    if (b) return wi * 32 + __ffs(b) - 1;
    if (++wi >= FINDER_WORDS) return -1;
    b = w[wi];
  }
}

// eachmatch: greedy leftmost non-overlapping matches -> 1-based offsets + add
__device__ __forceinline__ int greedy_list(const uint32_t* w, int mlen, int add, uint8_t* out) {
  int n = 0, p = 0;
  while ((p = next_bit(w, p)) >= 0) { out[n++] = (uint8_t)(p + 1 + add); p += mlen; }
  return n;
}

// seq: nucleotide codes 0..3 in shared memory.  Returns able; *lo (0-based nt offset) and *len3 (nt length) are set when a
// CDR3 was found (else *lo = -1).  All lanes return the same values.
__device__ inline bool finder_warp(const uint8_t* seq, int lgt, const FinderDev& fd, bool array_version, FinderScratch& fs,
                                   int lane, int* lo, int* len3) {
  bool able = false;
  *lo = -1; *len3 = 0;
  for (int shift = 0; shift < 3; shift++) {
    const int naa = (lgt - shift) / 3;
    if (naa <= 0) continue;
    for (int p = lane; p < naa; p += 32) {
      const uint8_t* s = seq + shift + 3 * p;
      const char ch = FINDER_CODON[s[0] * 16 + s[1] * 4 + s[2]];
      fs.aa[p] = ch == '*' ? 26 : (uint8_t)(ch - 'A');
    }
    __syncwarp();
    bool anyC = motif_scan(fd.m[M_C], fs.aa, naa, fs.cw, lane);
    bool anyF = motif_scan(fd.m[M_F], fs.aa, naa, fs.fw, lane);
    int code = 0, cadd = fd.coffset, fadd = fd.foffset;
    int clen = fd.m[M_C].len, flen = fd.m[M_F].len;
    if (anyC != anyF) {
      code = 2;
      if (!anyC) {
        const MotifDev& im = fd.m[array_version ? M_INNER_C : M_HARD_C];
        anyC = motif_scan(im, fs.aa, naa, fs.cw, lane);
        cadd = 0; clen = im.len;
      } else {
        const MotifDev& im = fd.m[array_version ? M_INNER_F : M_HARD_F];
        anyF = motif_scan(im, fs.aa, naa, fs.fw, lane);
        fadd = 2; flen = im.len;
      }
    }
    int r_lo = -1, r_hi = -1;
    if (anyC && anyF) {
      if (lane == 0) {
        const int nC = greedy_list(fs.cw, clen, cadd, fs.Cx);
        const int nF = greedy_list(fs.fw, flen, fadd, fs.Fx);
        for (int ci = nC - 1; ci >= 0; ci--) {            // Cx reversed: descending
          const int xc = fs.Cx[ci];
          for (int fi = 0; fi < nF; fi++) {
            const int xf = fs.Fx[fi], d = xf - xc;
            if (d < 6 || d > 34) continue;
            if (ci > 0) { const int d2 = xf - fs.Cx[ci - 1]; if (d2 >= 7 && d2 <= 35) continue; }
            bool stop = false;
            for (int k = xc; k <= xf && !stop; k++) stop = (k >= 1 && k <= naa && fs.aa[k - 1] == 26);
            if (stop) continue;
            r_lo = (xc - 1) * 3; r_hi = xf * 3;            // 0-based half-open within the frame
            break;
          }
        }
      }
      r_lo = __shfl_sync(0xffffffffu, r_lo, 0);
      r_hi = __shfl_sync(0xffffffffu, r_hi, 0);
      if (r_lo >= 0) code = 3;
    }
    __syncwarp();
    able = able || code > 0;
    if (code == 3) { *lo = r_lo + shift; *len3 = r_hi - r_lo; break; }
  }
  return able;
}

}  // namespace catt
