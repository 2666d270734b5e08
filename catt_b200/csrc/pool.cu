// pool.cu — K6/K7/K8: the k-mer table ("pool"), the greedy extension of partial reads with finder!(array), and the
// D-segment global alignment.
//
// Replaces convert2Dict / depcature / bbk / first_serveral / last_serveral (catt.jl:426-497, 544-548), ratio,
// fillpo_*!, extend_right!, extend_left! (catt.jl:341-424), the array finder! call sites (catt.jl:552-577) and
// selBestMatchD (Jtool.jl:40-52).
//
// The reference's table is NOT a histogram: convert2Dict stores res[kmer] = position (1..11) and per-read dicts are
// merged last-writer-wins in list order (Vpart, then Jpart overriding).  On the GPU every (read, position) does
// atomicMax(slot, order << 8 | position) into an open-addressing hash table in HBM, which yields exactly the
// sequential last-writer value; pool['G'^k] = 0 (catt.jl:545-548) is applied at lookup.
#include "finder.cuh"
#include "internal.cuh"

namespace catt {
namespace {

constexpr unsigned FULL = 0xffffffffu;
constexpr int PW = 4;   // warps per CTA
constexpr unsigned long long EMPTY = ~0ull;

__device__ __forceinline__ uint32_t sam_code(const uint32_t* __restrict__ w, const uint32_t* __restrict__ nm, int len, int i, int strand) {
  const int p = strand ? len - 1 - i : i;
  if ((nm[p >> 5] >> (p & 31)) & 1u) return 4u;
  const uint32_t c = (w[p >> 4] >> ((p & 15) * 2)) & 3u;
  return strand ? 3u - c : c;
}

__device__ __forceinline__ uint64_t mix64(uint64_t x) {      // table slot hash (murmur3 finaliser)
  x ^= x >> 33; x *= 0xff51afd7ed558ccdull; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ull; x ^= x >> 33;
  return x;
}

// Table addressing.  The table is `nsub` equal sub-tables laid back to back; a key lives in sub-table hash % nsub (the GPU that
// owns the key when one sample is spread over several GPUs, shard.cu) at slot umulhi(hash, sub_slots) (no power-of-two rounding:
// the load factor stays at 1/2 whatever the sample size), linear probing inside the sub-table.  One GPU: nsub == 1.
__device__ __forceinline__ uint32_t table_sub(const KmerTable& t, uint64_t h) { return t.nsub == 1 ? 0u : (uint32_t)(h & 0xffffffffu) % t.nsub; }
__device__ __forceinline__ void table_put_h(const KmerTable& t, uint64_t key, uint64_t h, unsigned long long val) {
  const uint64_t base = (uint64_t)table_sub(t, h) * t.sub_slots;
  uint64_t slot = __umul64hi(h, t.sub_slots);
  while (true) {
    unsigned long long* p = reinterpret_cast<unsigned long long*>(t.slots + base + slot);
    const unsigned long long prev = atomicCAS(p, EMPTY, (unsigned long long)key);
    if (prev == EMPTY || prev == key) { atomicMin(p + 1, ~val); return; }
    if (++slot == t.sub_slots) slot = 0;
  }
}
// Warp-aggregated insert: lanes of the warp that hold the same key (reads of one clonotype share their k-mers) elect one lane,
// which writes the largest (order, position) of the group - one atomic pair per distinct key instead of one per lane.
__device__ __forceinline__ void table_put_warp(const KmerTable& t, bool active, uint64_t key, unsigned long long val) {
  const unsigned live = __ballot_sync(FULL, active);
  if (!active) return;
  const unsigned peers = __match_any_sync(live, (unsigned long long)key);
  unsigned long long best = val;
  if (peers & (peers - 1)) {                                   // the group has more than one lane (rare): its maximum, by halves
    const unsigned hi = __reduce_max_sync(peers, (unsigned)(val >> 32));
    const unsigned lo = __reduce_max_sync(peers, (unsigned)(val >> 32) == hi ? (unsigned)val : 0u);
    best = ((unsigned long long)hi << 32) | lo;
  }
  if ((int)(threadIdx.x & 31) == __ffs(peers) - 1) table_put_h(t, key, mix64(key), best);
}

__device__ __forceinline__ int table_get(const KmerTable& t, uint64_t key, uint64_t gkey) {
  if (key == gkey) return 0;
  const uint64_t h = mix64(key);
  const uint64_t base = (uint64_t)table_sub(t, h) * t.sub_slots;
  uint64_t slot = __umul64hi(h, t.sub_slots);
  while (true) {
    const ulonglong2 s = __ldg(t.slots + base + slot);        // the table is complete and read-only by now
    if (s.x == EMPTY) return 0;
    if (s.x == key) return (int)(~s.y & 0xff);
    if (++slot == t.sub_slots) slot = 0;
  }
}

// The k-mer at SAM position `pos` of a read WITHOUT N (reads with N never reach Vpart/Jpart, catt.jl:103,124,299), first base in
// the most significant 2-bit group, straight from the packed words: one 2k-bit window, O(1) per k-mer instead of k base reads.
// Forward strand: the window holds base pos in its lowest group -> reverse the groups.  Reverse strand: SAM base i is the
// complement of read base len-1-i, so the window [len-pos-k, len-pos) already has the first SAM base on top -> complement only.
__device__ __forceinline__ uint64_t kmer_at(const uint32_t* __restrict__ w, int len, int pos, int k, int strand) {
  const int start = strand ? len - pos - k : pos;
  const int o = 2 * start, wi = o >> 5, s = o & 31;
  const uint64_t lo = (uint64_t)w[wi] | ((uint64_t)w[wi + 1] << 32);
  uint64_t win = lo >> s;
  if (s) win |= (uint64_t)w[wi + 2] << (64 - s);              // (the N-mask words and the buffer's padding follow: in bounds)
  const uint64_t kmask = (1ull << (2 * k)) - 1;               // k <= 31
  if (strand) return ~win & kmask;
  uint64_t r = __brevll(win & kmask);                          // bit reversal, then swap the two bits of every group back
  r = ((r & 0x5555555555555555ull) << 1) | ((r >> 1) & 0x5555555555555555ull);
  return r >> (64 - 2 * k);
}

// (read, position) pair `t` of the pool: item t / 11, position t % 11 of its k+10 nt fragment; every lane of the warp owns one
__device__ __forceinline__ bool pool_entry(const ReadsDev& reads, const PoolItem* __restrict__ items, int64_t t, int64_t n_entries, int k,
                                           uint64_t* key, unsigned long long* val) {
  if (t >= n_entries) return false;
  const PoolItem pi = items[t / 11];
  const int lane = (int)(t % 11);
  const int f = min((int)pi.seq_len, k + 10);
  if (lane >= f - k + 1) return false;
  const int foff = pi.seq_off + (pi.side == 0 ? pi.seq_len - f : 0);     // last_serveral / first_serveral
  *key = kmer_at(reads.words + reads.off[pi.read], reads.len[pi.read], foff + lane, k, pi.strand);
  *val = ((unsigned long long)pi.order << 8) | (unsigned)(lane + 1);
  return true;
}

// every (read, position): one lane, one window, one warp-aggregated insert
__global__ void __launch_bounds__(PW * 32) pool_kernel(ReadsDev reads, const PoolItem* __restrict__ items, int64_t n, int k, KmerTable tab) {
  const int64_t n_entries = n * 11;
  for (int64_t t0 = ((int64_t)blockIdx.x * PW + (threadIdx.x >> 5)) * 32; t0 < n_entries; t0 += (int64_t)gridDim.x * PW * 32) {
    uint64_t key = 0; unsigned long long val = 0;
    const bool ok = pool_entry(reads, items, t0 + (threadIdx.x & 31), n_entries, k, &key, &val);
    table_put_warp(tab, ok, key, val);
  }
}

// ---- one sample over several GPUs (shard.cu): the (k-mer, value) pairs of this GPU's reads, bucketed by the GPU that owns the
//      key; the owners insert what they receive (pool_insert_kernel) and the sub-tables are all-gathered.
//      pass 0 counts per owner, pass 1 scatters (cursor = exclusive offsets of the buckets, advanced by warp-aggregated atomics)
template <int PASS>
__global__ void __launch_bounds__(PW * 32) pool_emit_kernel(ReadsDev reads, const PoolItem* __restrict__ items, int64_t n, int k, KmerTable tab,
                                                            unsigned long long* __restrict__ cursor, PoolEntry* __restrict__ out) {
  const int64_t n_entries = n * 11;
  for (int64_t t0 = ((int64_t)blockIdx.x * PW + (threadIdx.x >> 5)) * 32; t0 < n_entries; t0 += (int64_t)gridDim.x * PW * 32) {
    uint64_t key = 0; unsigned long long val = 0;
    const bool ok = pool_entry(reads, items, t0 + (threadIdx.x & 31), n_entries, k, &key, &val);
    const unsigned live = __ballot_sync(FULL, ok);
    if (!ok) continue;
    const uint32_t dest = table_sub(tab, mix64(key));
    const unsigned peers = __match_any_sync(live, dest);
    const int leader = __ffs(peers) - 1, lane = threadIdx.x & 31;
    unsigned long long base = 0;
    if (lane == leader) base = atomicAdd(&cursor[dest], (unsigned long long)__popc(peers));
    base = __shfl_sync(peers, base, leader);
    if (PASS == 1) out[base + __popc(peers & ((1u << lane) - 1))] = PoolEntry{key, val};
  }
}
__global__ void __launch_bounds__(256) pool_insert_kernel(const PoolEntry* __restrict__ in, int64_t n, KmerTable tab) {
  for (int64_t t0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) & ~31ll; t0 < n; t0 += (int64_t)gridDim.x * blockDim.x) {
    const int64_t t = t0 + (threadIdx.x & 31);
    PoolEntry e{0, 0};
    if (t < n) e = in[t];
    table_put_warp(tab, t < n, e.key, e.val);
  }
}


__device__ __forceinline__ bool ratio_ok(int x, int y) { return x > y ? (x <= 10 * y) : (10 * x >= y); }   // catt.jl:341

// The greedy walk over the table (extend_right! / extend_left!, catt.jl:341-424) is a chain of dependent random lookups - one HBM
// round trip per added base, up to ~100 bases - so it is latency bound and wants as many chains in flight as the SM can hold:
// FOUR lanes per partial read (one candidate base each), eight reads per warp.  The start k-mer comes straight from the packed
// words (kmer_at); the added bases go to ExtendOut::ext2, 16 per word.  The extended reads then go through finder!(array) one
// warp per read (extend_finish_kernel).
__global__ void __launch_bounds__(PW * 32) extend_walk_kernel(ReadsDev reads, const PoolItem* __restrict__ items, const uint32_t* __restrict__ partial_idx,
                                                              int64_t n_partial, int k, KmerTable tab, ExtendOut* __restrict__ out) {
  const int lane = threadIdx.x & 31, c = lane & 3, grp = lane >> 2;
  const uint64_t kmask = (k >= 32) ? ~0ull : ((1ull << (2 * k)) - 1);
  uint64_t gkey = 0;
  for (int i = 0; i < k; i++) gkey = (gkey << 2) | 2;
  const int64_t ngroups = (int64_t)gridDim.x * PW * 8;
  for (int64_t it0 = ((int64_t)blockIdx.x * PW + (threadIdx.x >> 5)) * 8; it0 < n_partial; it0 += ngroups) {
    const int64_t it = it0 + grp;
    bool active = false;
    int side = 0, len0 = 0, cur = 0, n_ext = 0;
    uint64_t seg = 0;
    if (it < n_partial) {
      const PoolItem pi = items[partial_idx[it]];
      side = pi.side; len0 = pi.seq_len;
      if (len0 >= k) {
        const int s0 = side == 0 ? len0 - k : 0;
        seg = kmer_at(reads.words + reads.off[pi.read], reads.len[pi.read], pi.seq_off + s0, k, pi.strand);
        cur = table_get(tab, seg, gkey);                     // (the four lanes of the group read the same slot: one sector)
        active = 1 < 151 - len0;
      }
    }
    uint32_t wv = 0;                                         // the 16 bases being collected (meaningful on the group's first lane)
    while (__any_sync(FULL, active)) {
      // candidates in A,C,G,T order: right -> bw_neighbors(seg) = c + seg[1:k-1]; left -> fw_neighbors(seg) = seg[2:k] + c
      const uint64_t cand = side == 0 ? (((uint64_t)c << (2 * (k - 1))) | (seg >> 2)) : (((seg << 2) | (uint64_t)c) & kmask);
      int val = 0;
      if (active) val = table_get(tab, cand, gkey);
      // stable sort by value descending, then the first that passes == max value among passing, ties -> smallest c
      const bool pass = active && val != 0 && ratio_ok(val, cur) && cand != seg;
      int key = pass ? ((val << 2) | (3 - c)) : 0;
      key = max(key, __shfl_xor_sync(FULL, key, 1));
      key = max(key, __shfl_xor_sync(FULL, key, 2));
      if (active) {
        if (key) {
          const int cw = 3 - (key & 3);
          seg = side == 0 ? (((uint64_t)cw << (2 * (k - 1))) | (seg >> 2)) : (((seg << 2) | (uint64_t)cw) & kmask);
          cur = key >> 2;
          wv |= (uint32_t)cw << (2 * (n_ext & 15));
          n_ext++;
          if ((n_ext & 15) == 0) { if (c == 0) out[it].ext2[(n_ext >> 4) - 1] = wv; wv = 0; }
          active = n_ext + 1 < 151 - len0;                   // idx = n_ext + 1 in the reference's loop
        } else {
          active = false;
        }
      }
    }
    if (it < n_partial && c == 0) {
      ExtendOut* o = &out[it];
      int w = n_ext >> 4;
      if (n_ext & 15) o->ext2[w++] = wv;
      for (; w < EXT_WORDS; w++) o->ext2[w] = 0;
      o->n_ext = (int16_t)n_ext;
    }
  }
}

struct ExtWarp {
  uint8_t seq[MAX_READ + 160 + EXT_MAX + 8];
  FinderScratch fs;
};

// one warp per partial read: the extended read (right: seq * ext, left: reverse(ext) * seq), then finder!(array) on it
__global__ void __launch_bounds__(PW * 32) extend_finish_kernel(ReadsDev reads, const FinderDev* __restrict__ gfd, const PoolItem* __restrict__ items,
                                                                const uint32_t* __restrict__ partial_idx, int64_t n_partial,
                                                                ExtendOut* __restrict__ out, uint32_t* __restrict__ rescued) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  FinderDev* fd = reinterpret_cast<FinderDev*>(smem_raw);
  ExtWarp* wsall = reinterpret_cast<ExtWarp*>(smem_raw + sizeof(FinderDev));
  for (int i = threadIdx.x; i < (int)(sizeof(FinderDev) / 4); i += blockDim.x)
    reinterpret_cast<uint32_t*>(fd)[i] = reinterpret_cast<const uint32_t*>(gfd)[i];
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  ExtWarp& ws = wsall[warp];
  for (int64_t it = (int64_t)blockIdx.x * PW + warp; it < n_partial; it += (int64_t)gridDim.x * PW) {
    const PoolItem pi = items[partial_idx[it]];
    const int m = reads.len[pi.read];
    const uint32_t* __restrict__ w = reads.words + reads.off[pi.read];
    const uint32_t* __restrict__ nm = w + read_base_words(m);
    const int len0 = pi.seq_len;
    ExtendOut* o = &out[it];
    const int n_ext = o->n_ext;
    // the slice sits at offset 152 so that a left extension can be prepended in place
    uint8_t* base = ws.seq + 152;
    for (int i = lane; i < len0; i += 32) base[i] = (uint8_t)sam_code(w, nm, m, pi.seq_off + i, pi.strand);
    uint8_t* ext_seq = base;
    if (pi.side == 0) {
      for (int i = lane; i < n_ext; i += 32) base[len0 + i] = (uint8_t)((o->ext2[i >> 4] >> (2 * (i & 15))) & 3u);
    } else {
      ext_seq = base - n_ext;
      for (int i = lane; i < n_ext; i += 32) { const int e = n_ext - 1 - i; ext_seq[i] = (uint8_t)((o->ext2[e >> 4] >> (2 * (e & 15))) & 3u); }
    }
    __syncwarp();
    int lo, l3;
    const bool able = finder_warp(ext_seq, len0 + n_ext, *fd, true, ws.fs, lane, &lo, &l3);
    if (lane == 0) {
      o->cdr3_off = (int16_t)lo; o->cdr3_len = (int16_t)l3; o->able = able;
      if (lo >= 0) atomicAdd(rescued, 1u);
    }
    __syncwarp();
  }
}

// selBestMatchD: global affine, match 5, mismatch -3, gap(k) = -(4+k); keep score+len(seq)-len(D)+4 > 35; first max wins
constexpr int D_MAX = 32;
__global__ void best_d_kernel(const char* __restrict__ cdr3, const int32_t* __restrict__ off, int64_t n, const char* __restrict__ dseq,
                              const int32_t* __restrict__ doff, int nd, int32_t* __restrict__ out) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n) return;
  const char* a = cdr3 + off[t];
  const int m = off[t + 1] - off[t];
  int best = -1, best_score = 0;
  for (int d = 0; d < nd; d++) {
    const char* b = dseq + doff[d];
    const int nb = doff[d + 1] - doff[d];
    int H[D_MAX + 1], E[D_MAX + 1];
    const int NEG = -1000000;
    H[0] = 0; E[0] = NEG;
    for (int j = 1; j <= nb; j++) { H[j] = -(4 + j); E[j] = NEG; }
    for (int i = 1; i <= m; i++) {
      int hdiag = H[0], f = NEG;
      H[0] = -(4 + i);
      const uint8_t ca = nt_code(a[i - 1]);
      for (int j = 1; j <= nb; j++) {
        const int e = max(H[j] - 5, E[j] - 1);
        f = max(H[j - 1] - 5, f - 1);
        const int s = (ca == nt_code(b[j - 1]) && ca < 4) ? 5 : -3;
        const int h = max(hdiag + s, max(e, f));
        hdiag = H[j]; H[j] = h; E[j] = e;
      }
    }
    const int sc = H[nb];
    if (!(sc + m - nb + 4 > 35)) continue;
    if (best < 0 || sc > best_score) { best = d; best_score = sc; }
  }
  out[t] = best;
}

}  // namespace

int launch_pool(const ReadsDev& reads, const PoolItem* d_items, int64_t n, int kmer, KmerTable tab, int sm_count, cudaStream_t st, int* launches) {
  if (n == 0) return CATT_OK;
  const int grid = (int)std::min<int64_t>((n * 11 + PW * 32 - 1) / (PW * 32), (int64_t)sm_count * 16);
  pool_kernel<<<grid, PW * 32, 0, st>>>(reads, d_items, n, kmer, tab);
  CATT_CUDA(cudaGetLastError());
  *launches += 1;
  return CATT_OK;
}

int launch_pool_emit(const ReadsDev& reads, const PoolItem* d_items, int64_t n, int kmer, KmerTable tab, int pass, unsigned long long* d_cursor,
                     PoolEntry* d_out, int sm_count, cudaStream_t st, int* launches) {
  if (n == 0) return CATT_OK;
  const int grid = (int)std::min<int64_t>((n * 11 + PW * 32 - 1) / (PW * 32), (int64_t)sm_count * 16);
  if (pass == 0) pool_emit_kernel<0><<<grid, PW * 32, 0, st>>>(reads, d_items, n, kmer, tab, d_cursor, d_out);
  else pool_emit_kernel<1><<<grid, PW * 32, 0, st>>>(reads, d_items, n, kmer, tab, d_cursor, d_out);
  CATT_CUDA(cudaGetLastError());
  *launches += 1;
  return CATT_OK;
}

int launch_pool_insert(const PoolEntry* d_in, int64_t n, KmerTable tab, int sm_count, cudaStream_t st, int* launches) {
  if (n == 0) return CATT_OK;
  const int grid = (int)std::min<int64_t>((n + 255) / 256, (int64_t)sm_count * 16);
  pool_insert_kernel<<<grid, 256, 0, st>>>(d_in, n, tab);
  CATT_CUDA(cudaGetLastError());
  *launches += 1;
  return CATT_OK;
}

int launch_extend(const ReadsDev& reads, const FinderDev* d_finder, const PoolItem* d_items, const uint32_t* d_partial_idx,
                  int64_t n_partial, int kmer, KmerTable tab, ExtendOut* d_out, uint32_t* d_rescued, int sm_count, cudaStream_t st, int* launches) {
  if (n_partial == 0) return CATT_OK;
  const int gridw = (int)std::min<int64_t>((n_partial + PW * 8 - 1) / (PW * 8), (int64_t)sm_count * 16);
  extend_walk_kernel<<<gridw, PW * 32, 0, st>>>(reads, d_items, d_partial_idx, n_partial, kmer, tab, d_out);
  CATT_CUDA(cudaGetLastError());
  const size_t smem = sizeof(FinderDev) + sizeof(ExtWarp) * PW;
  static std::atomic<unsigned long long> optin{0};
  int rc = smem_optin(extend_finish_kernel, (int)smem, optin);
  if (rc) return rc;
  const int grid = (int)std::min<int64_t>((n_partial + PW - 1) / PW, (int64_t)sm_count * 8);
  extend_finish_kernel<<<grid, PW * 32, smem, st>>>(reads, d_finder, d_items, d_partial_idx, n_partial, d_out, d_rescued);
  CATT_CUDA(cudaGetLastError());
  *launches += 2;
  return CATT_OK;
}

int launch_best_d(const char* d_cdr3, const int32_t* d_off, int64_t n, const char* d_dseq, const int32_t* d_doff, int nd,
                  int32_t* d_out, cudaStream_t st) {
  if (n == 0) return CATT_OK;
  best_d_kernel<<<(unsigned)((n + 127) / 128), 128, 0, st>>>(d_cdr3, d_off, n, d_dseq, d_doff, nd, d_out);
  CATT_CUDA(cudaGetLastError());
  return CATT_OK;
}

}  // namespace catt
