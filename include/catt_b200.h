/* catt_b200.h — C ABI of libcatt_b200.so, the B200-native replacement for CATT's per-read CDR3 extraction
 * hot path.  Plain pointers and sizes only; no C++/torch types cross this boundary.
 *
 * The reference (GuoBioinfoLab/CATT, Julia) has no FFI for this path: today's boundary is a subprocess + SAM
 * file (catt.jl:43-54) consumed by Julia functions.  Each entry point below cites the reference code it
 * replaces; INTEGRATION.md shows the `ccall` stubs a maintainer adds to catt.jl.
 *
 * Conventions: every function returns 0 on success or a negative CATT_E_* code; catt_last_error() returns a
 * thread-local message.  No exception crosses the ABI.  The library owns every buffer it returns until
 * catt_result_free().  One context per (process, GPU) - or one context for a group of GPUs (catt_config.ngpu); calls on a
 * context are blocking and must not overlap.
 * There is NO CPU fallback: without a usable CUDA device catt_ctx_create fails with CATT_E_CUDA.
 */
#ifndef CATT_B200_H
#define CATT_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CATT_OK 0
#define CATT_E_ARG (-1)      /* bad argument / unsupported configuration (cf. catt.jl:712-722) */
#define CATT_E_IO (-2)       /* unreadable FASTA/FASTQ (reference: Julia SystemError) */
#define CATT_E_CUDA (-3)     /* no device, launch failure, out of device memory */
#define CATT_E_LIMIT (-4)    /* input exceeds a compiled-in limit (read longer than CATT_MAX_READ_LEN, ...) */

#define CATT_MAX_READ_LEN 512
#define CATT_KMER_AUTO 10000 /* catt:18, catt.jl:156 */

typedef struct catt_ctx catt_ctx;
typedef struct catt_result catt_result;

/* Everything read_alignrs()/catt() pull out of parsed_args and refg (catt.jl:187-205, 507-517; reference.jl). */
typedef struct catt_config {
  const char* v_fasta;      /* refg[..]["vregion"]; a sibling <fasta>.pac (bwa index) is used for N bases if present */
  const char* j_fasta;      /* refg[..]["jregion"] */
  const char* cmotif;       /* regex text, reference.jl */
  const char* fmotif;
  const char* inner_c;
  const char* inner_f;
  int32_t coffset;
  int32_t foffset;
  int32_t min_score;        /* args["bowsc"] -> bwa mem -T (catt.jl:43,145); default 12 */
  int32_t kmer;             /* args["kmer"]; CATT_KMER_AUTO = adaptive (catt.jl:156-172) */
  int32_t julia_nthreads;   /* Threads.nthreads(): fixes the record order of catt.jl:228,256-276 */
  int32_t allow_v;          /* real_score allowance in assignV (catt.jl:100): 3 at HEAD, 0 reproduces the bundled golden */
  int32_t allow_j;          /* real_score allowance in assignJ (catt.jl:121): 9 */
  int32_t device;           /* CUDA device ordinal */
  int32_t n_d;              /* refg[..]["Dregion"] length (0 = chain has no D segments) */
  const char* const* d_names;
  const char* const* d_seqs;
  int32_t keep_records;     /* 1: every result also carries the per-read alignment records (catt_align_records); costs a
                               64 B/read device->host copy, so it is off on the production path */
  int32_t chunk_reads;      /* Stage A batch size of the pipelined path (0 = default 1048576): host packing + H2D of chunk
                               i+1 run under the kernels of chunk i */
  int32_t only_cdr3;        /* 0: Vpart / Jpart exactly as catt() holds them before catt.jl:581 (every read, cdr3 "None" included);
                               1: as they stand after catt.jl:581-582, `filter!(x -> x.cdr3 != "None", ...)` - the only reads the
                               rest of catt() ever touches; catt_info still carries the counts of the full lists for the log
                               lines.  Cuts the device->host traffic ~50x on typical samples. */
  int32_t ngpu;             /* <= 1: one GPU (`device`).  N > 1: ONE sample is spread over N GPUs of this process: the reads are
                               split into N contiguous blocks, seeding / SW / assignV,J / finder! / extension run on every block's GPU,
                               and the global steps of the reference - the SAM order and QNAME dedup (catt.jl:45-48), share_name and
                               the both-mapped pairing (catt.jl:237, 288-319), the last-writer-wins k-mer table (catt.jl:426-497,
                               544-548) and the final lists - are merged over NCCL / NVLink.  The calls and the results are the same
                               as with one GPU (bit-identical lists); catt_info.ms_coll / bytes_coll report the exchange. */
  const int32_t* devices;   /* ngpu CUDA device ordinals; NULL = 0 .. ngpu-1.  Repeated ordinals are allowed (several ranks on one
                               GPU, exchanging through peer copies instead of NCCL): that is how the multi-rank path is tested on
                               a single-GPU box. */
  int32_t clonotypes;       /* 1: every result also carries its clonotype table (catt_result_clonotypes): distinct (CDR3, V, J) with
                               read count and merged quality string, built on the device from the lists - what nbsorb's grouping
                               (Jtool.jl:475-490) and the output counter (catt.jl:644) compute from the reads */
  int32_t reserved_;
} catt_config;

/* One SAM-visible alignment record = the line `bwa mem | samtools ... | awk` would leave for one read against one
 * reference set (catt.jl:43-51), reduced to the fields the Julia code reads (Jtool.jl:76-124, catt.jl:69-132). */
typedef struct catt_aln {
  int32_t ordinal;    /* 0-based read index in the input */
  int32_t rid;        /* reference record index (RNAME), -1 if unmapped */
  int16_t strand;     /* 1 = read was reverse-complemented (FLAG 16) */
  int16_t mapped;     /* 1 if AS >= min_score */
  int16_t as;         /* AS:i */
  int16_t qb;         /* soft clip before the alignment; GetStartCigar = qb+1 */
  int16_t m_end;      /* GetEndCigar: 1-based end of the first M block */
  int16_t rb;         /* POS-1 (record-local) */
  int16_t qe;         /* alignment end on the read (exclusive) */
  int16_t re;         /* alignment end on the record (exclusive) */
  int16_t nm;         /* NM:i */
  int16_t mi;         /* sum of M and I lengths ("bases mapped (cigar)") */
  int16_t ncand;      /* (record,strand) candidates scored for this read */
  int16_t ntied;      /* candidates sharing the best score */
} catt_aln;

/* Myread (Jtool.jl:12-19) after finder!/extension. */
typedef struct catt_read {
  const char* seq;
  const char* qual;
  int32_t seq_len;
  int32_t qual_len;     /* J-side reads carry a qual string of a different length (catt.jl:125-126) */
  int32_t v_id;         /* index into the V reference names, -1 = "None" */
  int32_t j_id;
  int32_t cdr3_off;     /* 0-based offset of cdr3 in seq, -1 = "None" */
  int32_t cdr3_len;
  int32_t ordinal;      /* provenance: input read index */
  uint8_t able;
  uint8_t flags;        /* 1 V part read, 2 J part read, 4 both-mapped, 8 qual slice clipped (reference would throw) */
  uint8_t pad_[2];
} catt_read;

/* One row of the clonotype table: all returned reads with the same CDR3 nucleotides, V call and J call.  Rows are ordered by
 * (cdr3_len, cdr3, v_id, j_id) - within a length that is the order of nbsorb's `sort!(val, by = x -> x.cdr3)` (Jtool.jl:481) -
 * so the rows of one nbsorb call (one CDR3 length, catt.jl:624-633) are contiguous and the rows of one cdr3 adjacent:
 *   nbsorb's (SeqIns(cdr3, reduce(MQ, quals)), count)  =  per cdr3: per-position max of `qual`, sum of `count` over its rows;
 *   catt.jl:644's counter((vs, translate(cdr3), js, cdr3)) = the rows themselves. */
typedef struct catt_clonotype {
  const char* cdr3;       /* NUL-terminated nucleotides */
  const char* qual;       /* NUL-terminated; per-position maximum of the members' quality characters = reduce(MQ, ...) (Jtool.jl:466-468,487) */
  int32_t cdr3_len;
  int32_t v_id;           /* index into the V reference names, -1 = "None" */
  int32_t j_id;
  int32_t reserved_;
  int64_t count;          /* reads */
} catt_clonotype;

/* The numbers the reference logs (catt.jl:331,503-524,580) plus work counters. */
typedef struct catt_info {
  int32_t kmer;               /* args["kmer"] after get_kmer_fromsam */
  int32_t reserved_;
  double the_err;             /* global the_ERR after get_err_fromsam (already %e-rounded) */
  double avg_len;
  int64_t n_reads, v_hits, j_hits, v_final, j_final, share, pair_seq_mismatch, full_pushed;
  int64_t vpart, jpart, direct, left, rescued;
  int64_t cand_v, cand_j;     /* scored (read, record, strand) pairs */
  int64_t cells_v, cells_j;   /* DP cell updates: sum len(read) x len(record) */
  int64_t gapped_v, gapped_j; /* winners that needed the traceback kernel */
  int64_t kernel_launches;    /* CUDA kernels launched for this result */
  double ms_seed, ms_sw, ms_select, ms_extract, ms_pool, ms_h2d, ms_d2h, ms_host;  /* device/host stage times; seed / sw /
                                 select are those of the V reference set (the J set runs concurrently, see ms_j_stream) */
  int64_t cells_computed_v, cells_computed_j; /* DP cells actually computed: byte-identical records are scored once */
  double ms_order;            /* device time of the SAM ordering / QNAME dedup / pairing kernels (samtools sort | tac | awk) */
  double ms_pack;             /* host time packing reads that was NOT hidden under kernels */
  double ms_seed_kernel;      /* seed_kernel alone (ms_seed also holds the task scan / expansion) */
  double ms_pool_kernel;      /* pool_kernel alone (ms_pool also holds the extension kernel) */
  double ms_strings;          /* the kernel that writes the Myread strings into the result arenas */
  int64_t bytes_h2d, bytes_d2h; /* bytes the library copied host->device / device->host for this result */
  double ms_j_stream;         /* Stage A of the J reference set (seed + sw + select), on its own stream under the V kernels */
  /* ---- one sample over several GPUs (catt_config.ngpu > 1 or catt_ctx_join); all zero on one GPU.  Counters above are sums over
   *      the ranks, stage times the maximum over the ranks. */
  int32_t n_ranks;            /* GPUs that shared this sample */
  int32_t reserved2_;
  double ms_comm;             /* device time inside collectives (maximum over the ranks) = sum of ms_coll */
  int64_t bytes_comm;         /* bytes received over NVLink, summed over the ranks = sum of bytes_coll */
  double ms_coll[8];          /* per collective, CATT_COLL_* below */
  int64_t bytes_coll[8];
  double ms_stage_a_max, ms_stage_a_min;   /* wall time of the slowest / fastest rank's Stage A (host clock around its pipeline) */
} catt_info;

/* indices of catt_info.ms_coll / bytes_coll */
#define CATT_COLL_RECORDS 0     /* all-gather of the mapped reads' alignment records + QNAMEs (global SAM order / dedup / pairing) */
#define CATT_COLL_MATES 1       /* all-reduce of the packed reads of both-mapped pairs whose records sit on two GPUs */
#define CATT_COLL_FLAGS 2       /* all-reduce of one byte per work item: kept / list / partial -> positions in [Vpart; Jpart] */
#define CATT_COLL_KMER_A2A 3    /* all-to-all of (k-mer, order << 8 | position) pairs to the GPU that owns the key */
#define CATT_COLL_KMER_TABLE 4  /* all-gather of the merged k-mer sub-tables (the table is replicated for the extension) */
#define CATT_COLL_SIZES 5       /* all-reduce of the string sizes of the returned reads -> global arena offsets */
#define CATT_COLL_RESULT 6      /* sum-reduce of the result arenas to the root */

/* ---- lifecycle --------------------------------------------------------------------------------------------- */
/* Loads the V/J references (what `bwa index` + FASTA.Reader do, catt.jl:198-205) and builds the device tables. */
int catt_ctx_create(catt_ctx** out, const catt_config* cfg);
void catt_ctx_destroy(catt_ctx* ctx);
const char* catt_last_error(void);
const char* catt_version(void);

/* Host threads the library uses for packing reads and materialising the result strings (process-wide; 0 = all hardware
 * threads).  One process per GPU: give each rank its share of the cores (launchers such as torchrun export OMP_NUM_THREADS=1). */
int catt_set_host_threads(int32_t n);

int catt_ref_count(const catt_ctx* ctx, int which /*0=V,1=J*/);
const char* catt_ref_name(const catt_ctx* ctx, int which, int32_t id);      /* FASTA.identifier, catt.jl:202 */
int catt_ref_seq(const catt_ctx* ctx, int which, int32_t id, char* out, int32_t cap); /* alignment bases; returns length */

/* ---- one sample over several GPUs, one PROCESS per GPU (torchrun / MPI style launchers) ------------------------------------------
 * The in-process form needs nothing but catt_config.ngpu.  With one process per GPU, rank 0 calls catt_comm_id and hands the
 * 128 bytes to the other ranks through the launcher's own channel; every rank then creates its context (ngpu <= 1, its own
 * device) and calls catt_ctx_join (collective: ncclCommInitRank).  catt_run_reads_shard is then called by EVERY rank with its
 * contiguous block [lo, lo + n_local) of the sample's reads (blocks in rank order, together covering 0 .. n_total);
 * file_start are the GLOBAL file boundaries (nfiles + 1 entries, 0 .. n_total) as in catt_run_reads_files.  Rank 0 receives
 * the lists; every rank receives the (identical) counters. */
int catt_comm_id(uint8_t id[128]);
int catt_ctx_join(catt_ctx* ctx, int32_t nranks, int32_t rank, const uint8_t id[128]);
int catt_run_reads_shard(catt_ctx* ctx, int64_t n_total, int64_t lo, int64_t n_local, const char* names, const int64_t* name_off,
                         const char* seqs, const int64_t* seq_off, const char* quals, int32_t nfiles, const int64_t* file_start,
                         catt_result** out);

/* File form: every rank opens the same plain 4-line FASTQ file and ingests ITS byte range (cut at record starts) on its own GPU - N PCIe
 * links and N shares of the host cores in parallel, nothing parsed on the host; the read counts are exchanged before Stage A because
 * bwa's tie-break hashes the global read ordinal.  (catt_run_file on an in-process group does the same; other formats go through
 * the host parser there, and are refused here.) */
int catt_run_file_shard(catt_ctx* ctx, const char* fastq_path, catt_result** out);

/* Resident form of catt_run_reads_shard, for throughput measurements with the inputs already in HBM (bench.py `value`):
 * catt_shard_upload copies this rank's block to the device once (packed reads, QNAMEs, quality strings); catt_shard_run - collective,
 * repeatable - runs seeding / SW / selection on the resident block and then the global Stage B.  Without catt_ctx_join the block
 * must be the whole sample (lo = 0, n_local = n_total) and the call is the one-GPU path. */
int catt_shard_upload(catt_ctx* ctx, int64_t n_total, int64_t lo, int64_t n_local, const char* names, const int64_t* name_off,
                      const char* seqs, const int64_t* seq_off, const char* quals, int32_t nfiles, const int64_t* file_start);
int catt_shard_run(catt_ctx* ctx, catt_result** out);

/* ---- the hot path ------------------------------------------------------------------------------------------- */
/* Replaces input_convert (catt.jl:134-151) + read_alignrs (catt.jl:186-339) + the k-mer pool / extension /
 * finder!(array) part of catt() (catt.jl:544-577) for one single-end FASTQ/FASTA(.gz) file. */
int catt_run_file(catt_ctx* ctx, const char* fastx_path, catt_result** out);

/* Same for reads already in host memory.  names/seqs/quals are concatenated, NOT NUL-terminated; *_off has n+1
 * entries.  quals may be NULL (FASTA input).  Names are the raw header tokens; `/1` `/2` trimming is done inside. */
int catt_run_reads(catt_ctx* ctx, int64_t n, const char* names, const int64_t* name_off, const char* seqs,
                   const int64_t* seq_off, const char* quals, catt_result** out);

/* Paired-end samples (--f1/--f2, catt.jl:738-745).  The reference aligns each mate file on its own (two `bwa mem` runs, read
 * ordinals restart at 0) and runs the read_alignrs loop once per file - SAM order, QNAME dedup, share_name and the both-mapped
 * pairing are per file, args["kmer"] comes from the first file whose mean mapped length reaches 50, the_ERR from the last -
 * and then ONE k-mer table / extension pass over the joined Vpart / Jpart lists.  `paths` in --f1, --f2 order. */
int catt_run_files(catt_ctx* ctx, int32_t nfiles, const char* const* paths, catt_result** out);
/* Same for reads in host memory: the reads of all files back to back, file f = reads [file_start[f], file_start[f+1]). */
int catt_run_reads_files(catt_ctx* ctx, int64_t n, const char* names, const int64_t* name_off, const char* seqs,
                         const int64_t* seq_off, const char* quals, int32_t nfiles, const int64_t* file_start, catt_result** out);

/* `--bam` input (catt.jl:756-764: `samtools fastq -N in.bam > tmp.fastq`, then the single-end path): BAM (BGZF or uncompressed)
 * or SAM text.  Secondary / supplementary records are dropped, reverse-strand records are restored to read orientation, "/1" "/2"
 * are appended to READ1 / READ2 names, a missing quality string becomes 'B's - what `samtools fastq -N` writes - and the reads
 * go through the same path as catt_run_reads without ever becoming FASTQ text. */
int catt_run_bam(catt_ctx* ctx, const char* path, catt_result** out);
/* The same conversion written out as FASTQ (needs no GPU and no context): a stand-in for the `samtools fastq -N` step itself. */
int catt_bam_to_fastq(const char* bam_path, const char* fastq_path);
/* How the library reads a (possibly compressed) input file, written back out: plain bytes (method 0), BGZF members inflated in
 * parallel (1), a single-member gzip stream inflated in parallel by speculative block decoding, verified by its CRC-32 (2), or one
 * zlib stream (3: multi-member or unusual files).  The reference reads .gz inputs through zlib inside bwa (kseq + gzread, one
 * thread); this is the same bytes, faster.  Needs no GPU and no context; used by the tests and for timing. */
int catt_inflate_file(const char* path, const char* out_path, int32_t* method);

int catt_result_info(const catt_result* r, catt_info* out);
/* Order-sensitive 64-bit digests of Vpart (out[0]) and Jpart (out[1]) - every catt_read field and both strings of every entry:
 * equal digests = identical lists (used to show that a sample gives the same result on 1, 2, 4 and 8 GPUs). */
int catt_result_digest(const catt_result* r, uint64_t out[2]);
/* The same idea over the lists as TEXT - seq, qual, V name, J name, cdr3 ("None" when absent), able - i.e. over what the Julia side
 * sees; independent of ids / offsets, so another implementation of the path can compute the same number (the test oracle does). */
int catt_result_text_digest(const catt_ctx* ctx, const catt_result* r, uint64_t out[2]);
/* Vpart (which=0) / Jpart (1) in the reference's list order, ready for catt.jl:579 onwards. */
int catt_result_reads(const catt_result* r, int which, const catt_read** reads, int64_t* n);
/* Per-read alignment records (input order, one per read; unmapped reads have mapped=0): parity/diagnostics.
 * Only present when the context was created with keep_records = 1 (otherwise *n = 0). */
int catt_align_records(const catt_result* r, int which, const catt_aln** recs, int64_t* n);
/* The clonotype table of the result (catt_config.clonotypes = 1; otherwise *n = 0). */
int catt_result_clonotypes(const catt_result* r, const catt_clonotype** rows, int64_t* n);
void catt_result_free(catt_result* r);

/* Replaces the selBestMatchD loop (catt.jl:676-683, Jtool.jl:40-52): for each CDR3 string the index of the best D
 * segment of the context's D list, or -1 for "None". */
int catt_best_d(catt_ctx* ctx, const char* const* cdr3, int64_t n, int32_t* d_index_out);

/* ---- the two O(n x m) scans of `nbsorb`; grouping, thresholds and tree assignment stay in Julia --- */
/* Replaces the Threads.@threads loop over linkRLG (Jtool.jl:524-531 calling Jtool.jl:396-421).  All sequences of one nbsorb call have
 * the same length `lgt` (catt.jl:624-633) and are passed back to back, n x lgt bytes over ACGTN, no terminators.
 * `uplimit` = root_up (Jtool.jl:515-519), n_roots x 7 doubles; leaf_count = the leaves' frequencies.  Per leaf: statu = number of roots r
 * with HammingDistanceG3(leaf, root_r, argmin|uplimit_r - count|; skip = 3), capped at 2, and root_index = 1-based index of the first
 * such root (1 if none), exactly what linkRLG returns. */
int catt_link_leaves(catt_ctx* ctx, int32_t lgt, const char* roots, int64_t n_roots, const double* uplimit, const char* leaves,
                     const int64_t* leaf_count, int64_t n_leaves, int32_t* statu, int32_t* root_index);
/* Replaces the inner `for confm in highconf` scan of Jtool.jl:553-558: flag[q] = 1 iff some target t has
 * HammingDistanceG3(query_q, target_t, upper; skip) (the reference uses upper = 3, skip = 3).  The err_cnt tests around it stay in Julia. */
int catt_near_any(catt_ctx* ctx, int32_t lgt, const char* queries, int64_t n_queries, const char* targets, int64_t n_targets,
                  int32_t upper, int32_t skip, uint8_t* flag);

/* ---- stage-level entry points (sharding, benchmarks, parity tests) --------------------------------------------- */
/* Stage A = map2align for a shard of reads: seeding + SW + hit selection.  `first_ordinal` is the global index of
 * seqs[0] (bwa's hash tie-break uses the global read ordinal).  Writes n records per reference set. */
int catt_align_shard(catt_ctx* ctx, int64_t n, const char* seqs, const int64_t* seq_off, int64_t first_ordinal,
                     catt_aln* v_out, catt_aln* j_out);

/* Device-side exchange for one sample over several GPUs (catt_b200/shard.py): after catt_align_shard(..., NULL, NULL) the shard's
 * records sit at the start of the context's record arrays; catt_records_device returns their device address (arrays grown to hold
 * `capacity` records, contents kept) so the ranks can all-gather them over NCCL / NVLink without a host round trip, and
 * catt_assemble_device runs Stage B on the n records the caller left there. */
int catt_records_device(catt_ctx* ctx, int which, int64_t capacity, uint64_t* device_ptr);
int catt_assemble_device(catt_ctx* ctx, int64_t n, const char* names, const int64_t* name_off, const char* seqs,
                         const int64_t* seq_off, const char* quals, catt_result** out);

/* ... and the packed reads: after catt_align_shard the shard's 2-bit packed reads sit in the context's read buffers
 * (catt_shard_counts says how many reads / words / the longest read); catt_reads_device returns the device addresses of the word,
 * offset and length arrays, grown (contents kept) to the given capacities, so the ranks can all-gather them as well; and
 * catt_assemble_resident runs Stage B on n reads + records that are ALL already in the context's buffers - the root re-reads
 * nothing from the host except the QNAMEs (and, with only_cdr3, the quality strings of the few surviving reads). */
int catt_shard_counts(const catt_ctx* ctx, int64_t* n_reads, int64_t* n_words, int32_t* max_len);
int catt_reads_device(catt_ctx* ctx, int64_t cap_reads, int64_t cap_words, uint64_t* words_ptr, uint64_t* off_ptr, uint64_t* len_ptr);
int catt_assemble_resident(catt_ctx* ctx, int64_t n, int32_t max_len, const char* names, const int64_t* name_off,
                           const int64_t* seq_off, const char* quals, catt_result** out);

/* Device-resident variant used by bench.py: upload once, run Stage A repeatedly on the resident batch, read back. */
typedef struct catt_batch catt_batch;
int catt_batch_upload(catt_ctx* ctx, int64_t n, const char* seqs, const int64_t* seq_off, int64_t first_ordinal,
                      catt_batch** out);
int catt_batch_align(catt_ctx* ctx, catt_batch* b);                       /* async on the context's stream */
int catt_batch_sync(catt_ctx* ctx, catt_batch* b, catt_info* stage_info); /* waits; fills counters and stage times */
int catt_batch_download(catt_ctx* ctx, catt_batch* b, catt_aln* v_out, catt_aln* j_out);
void catt_batch_free(catt_batch* b);
/* The CUDA stream the context launches on, as an integer handle (cudaStream_t), for event timing by callers. */
uint64_t catt_ctx_stream(const catt_ctx* ctx);

/* Stage B = everything after the SAM files exist, given ALL reads of the sample and their records (gathered from
 * the shards): ordering, dedup, V/J join, assignV/J, finder!, k-mer table, extension. */
int catt_assemble(catt_ctx* ctx, int64_t n, const char* names, const int64_t* name_off, const char* seqs,
                  const int64_t* seq_off, const char* quals, const catt_aln* v_recs, const catt_aln* j_recs,
                  catt_result** out);

/* Stage B on a resident, aligned batch.  The first call moves the QNAMEs and quality strings of the batch to HBM as well (they
 * are inputs of Stage B); later calls on the same batch must pass the same buffers and run with every input resident. */
int catt_batch_assemble(catt_ctx* ctx, catt_batch* b, const char* names, const int64_t* name_off, const char* seqs,
                        const int64_t* seq_off, const char* quals, catt_result** out);

/* Synthetic read generator of the benchmark workload: read i is drawn from splitmix64(seed+i) from the
 * context's V/J records; fixed length; seqs/quals hold n*length bytes; names ("s<i>", up to 24 bytes each) + n+1 offsets. */
int catt_synth_reads(const catt_ctx* ctx, int64_t n, int64_t first_index, int32_t length, uint64_t seed, double p_err,
                     int32_t with_n, char* seqs, char* quals, char* names, int64_t* name_off);

/* ---- parity-test hooks (thin wrappers over single kernels) ----------------------------------------------------- */
/* Seeding only: cand_out[n][2*nrec] bytes, 1 where (record,strand) survives bwa-mem's seed filters. */
int catt_seed_candidates(catt_ctx* ctx, int which, int64_t n, const char* seqs, const int64_t* seq_off, uint8_t* cand_out);
/* SW kernel only: score read i against record rid[i] on strand[i]; out[i] = {score, end_read, end_record}. */
int catt_sw_pairs(catt_ctx* ctx, int which, int64_t n, const char* seqs, const int64_t* seq_off, const int32_t* rid,
                  const int32_t* strand, int32_t* out3);
/* finder! on raw nucleotide strings: out[i] = {able, cdr3_off (0-based, -1 none), cdr3_len}. */
int catt_finder(catt_ctx* ctx, int array_version, int64_t n, const char* seqs, const int64_t* seq_off, int32_t* out3);
/* real_score (catt.jl:58-84) alone, one SAM record per read: record id / strand / soft clip qb / GetEndCigar m_end / POS-1 rb and the
 * allowance (3 in assignV, 9 in assignJ); out[i] = 1 when the record passes. */
int catt_real_score(catt_ctx* ctx, int which, int64_t n, const char* seqs, const int64_t* seq_off, const int32_t* rid, const int32_t* strand,
                    const int32_t* qb, const int32_t* m_end, const int32_t* rb, int32_t allowance, uint8_t* out);
/* The ordering stage alone (`samtools sort -t AS | view -F 2308 | tac | awk '!a[$1]++'`, share_name, the NR % nthreads chunk order, the
 * name-sorted pairing: catt.jl:45-48, 221-237, 256-276, 288-297) for the reads of ONE file: from QNAMEs, read lengths and per-read
 * records to the work items in processing order - items3[3t] = kind (0 V-only, 1 J-only, 2 both-mapped), read, second read (the J
 * record's read of a pair); room for 3 * (mapped V + mapped J) ints.  counts5 = final V records, final J records, V-only items,
 * pairs, J-only items. */
int catt_order_records(catt_ctx* ctx, int64_t n, const char* names, const int64_t* name_off, const int32_t* read_len, const catt_aln* v_recs,
                       const catt_aln* j_recs, int32_t* items3, int64_t* n_items, int64_t* counts5);
/* Integer-ALU peak microbenchmark used as the DP roofline denominator: returns packed-s16x2 DPX op/s. */
int catt_alu_peak(catt_ctx* ctx, double* dpx_ops_per_s, double* ms);

#ifdef __cplusplus
}
#endif
#endif /* CATT_B200_H */
