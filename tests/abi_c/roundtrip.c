/* roundtrip.c - a plain-C client of libcatt_b200.so: what a non-Python, non-Julia caller (or the reference maintainer's own ccall
 * stub) sees through include/catt_b200.h.  Built by __graft_entry__.build() with gcc -std=c11 (the header must
 * stay C), linked against the in-tree library.
 *
 *   roundtrip layout
 *       prints sizeof / offsetof of every struct of the header, one "name value" per line: tests/test_abi.py compares them with
 *       the ctypes mirror (catt_b200/_lib.py), so the two cannot drift apart silently.  Needs no GPU.
 *   roundtrip run <resource_dir> <fastq> [ngpu [distinct]]     (distinct: one GPU per rank, NCCL from a process without torch)
 *       fills catt_config BY VALUE for hs TRB (the fields of reference.jl:24-37), runs catt_run_file, walks the catt_read arrays
 *       the way catt.jl:579 would and prints the counters and an order-sensitive checksum of (seq, qual, V name, J name, cdr3)
 *       over both lists.  The GPU test compares this line with the same numbers obtained through ctypes.
 */
#include <stddef.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "catt_b200.h"

#define P(type, field) printf(#type "." #field " %zu\n", offsetof(type, field))

static unsigned long long fnv(unsigned long long h, const char* s, size_t n) {
  for (size_t i = 0; i < n; i++) { h ^= (unsigned char)s[i]; h *= 0x100000001b3ull; }
  return h;
}

static int layout(void) {
  printf("sizeof.catt_config %zu\n", sizeof(catt_config));
  P(catt_config, v_fasta); P(catt_config, j_fasta); P(catt_config, cmotif); P(catt_config, fmotif); P(catt_config, inner_c); P(catt_config, inner_f);
  P(catt_config, coffset); P(catt_config, foffset); P(catt_config, min_score); P(catt_config, kmer); P(catt_config, julia_nthreads);
  P(catt_config, allow_v); P(catt_config, allow_j); P(catt_config, device); P(catt_config, n_d); P(catt_config, d_names); P(catt_config, d_seqs);
  P(catt_config, keep_records); P(catt_config, chunk_reads); P(catt_config, only_cdr3); P(catt_config, ngpu); P(catt_config, devices); P(catt_config, clonotypes);
  printf("sizeof.catt_clonotype %zu\n", sizeof(catt_clonotype));
  P(catt_clonotype, cdr3); P(catt_clonotype, qual); P(catt_clonotype, cdr3_len); P(catt_clonotype, v_id); P(catt_clonotype, j_id); P(catt_clonotype, count);
  printf("sizeof.catt_aln %zu\n", sizeof(catt_aln));
  P(catt_aln, ordinal); P(catt_aln, rid); P(catt_aln, strand); P(catt_aln, mapped); P(catt_aln, as); P(catt_aln, qb); P(catt_aln, m_end);
  P(catt_aln, rb); P(catt_aln, qe); P(catt_aln, re); P(catt_aln, nm); P(catt_aln, mi); P(catt_aln, ncand); P(catt_aln, ntied);
  printf("sizeof.catt_read %zu\n", sizeof(catt_read));
  P(catt_read, seq); P(catt_read, qual); P(catt_read, seq_len); P(catt_read, qual_len); P(catt_read, v_id); P(catt_read, j_id);
  P(catt_read, cdr3_off); P(catt_read, cdr3_len); P(catt_read, ordinal); P(catt_read, able); P(catt_read, flags);
  printf("sizeof.catt_info %zu\n", sizeof(catt_info));
  P(catt_info, kmer); P(catt_info, the_err); P(catt_info, avg_len); P(catt_info, n_reads); P(catt_info, v_hits); P(catt_info, j_hits);
  P(catt_info, v_final); P(catt_info, j_final); P(catt_info, share); P(catt_info, pair_seq_mismatch); P(catt_info, full_pushed);
  P(catt_info, vpart); P(catt_info, jpart); P(catt_info, direct); P(catt_info, left); P(catt_info, rescued); P(catt_info, cand_v);
  P(catt_info, cand_j); P(catt_info, cells_v); P(catt_info, cells_j); P(catt_info, gapped_v); P(catt_info, gapped_j);
  P(catt_info, kernel_launches); P(catt_info, ms_seed); P(catt_info, ms_sw); P(catt_info, ms_select); P(catt_info, ms_extract);
  P(catt_info, ms_pool); P(catt_info, ms_h2d); P(catt_info, ms_d2h); P(catt_info, ms_host); P(catt_info, cells_computed_v);
  P(catt_info, cells_computed_j); P(catt_info, ms_order); P(catt_info, ms_pack); P(catt_info, ms_seed_kernel); P(catt_info, ms_pool_kernel);
  P(catt_info, ms_strings); P(catt_info, bytes_h2d); P(catt_info, bytes_d2h); P(catt_info, ms_j_stream); P(catt_info, n_ranks);
  P(catt_info, ms_comm); P(catt_info, bytes_comm); P(catt_info, ms_coll); P(catt_info, bytes_coll); P(catt_info, ms_stage_a_max);
  P(catt_info, ms_stage_a_min);
  return 0;
}

static int run(const char* res, const char* fastq, int ngpu, int distinct) {
  char v[4096], j[4096];
  snprintf(v, sizeof v, "%s/TR/hs/TRB/TRBV-gai6-DNA.fa", res);
  snprintf(j, sizeof j, "%s/TR/hs/TRB/TRBJ-gai3-DNA.fa", res);
  /* reference.jl:24-37 (hs TRB CDR3) */
  static const char* d_names[3] = {"TRBD1*01", "TRBD2*01", "TRBD2*02"};
  static const char* d_seqs[3] = {"gggacagggggc", "gggactagcggggggg", "gggactagcgggaggg"};
  int32_t devices[16] = {0};
  catt_config cfg;
  memset(&cfg, 0, sizeof cfg);
  cfg.v_fasta = v; cfg.j_fasta = j;
  cfg.cmotif = "(LR|YF|YI|YL|YQ|YR)C(A|S|T|V|G|R|P|D)"; cfg.fmotif = "[FTYH]{1}FG[ADENPQS]{1}G";
  cfg.inner_c = "(CAS|CSA|CAW|CAT|CSV|CAI|CAR)"; cfg.inner_f = "place_holder";
  cfg.coffset = 2; cfg.foffset = 1; cfg.min_score = 12; cfg.kmer = CATT_KMER_AUTO; cfg.julia_nthreads = 2;
  cfg.allow_v = 3; cfg.allow_j = 9; cfg.device = 0; cfg.n_d = 3; cfg.d_names = d_names; cfg.d_seqs = d_seqs;
  cfg.only_cdr3 = 1; cfg.clonotypes = 1; cfg.ngpu = ngpu;
  cfg.devices = (ngpu > 1 && !distinct) ? devices : NULL;     /* all ranks on device 0 (peer copies), or devices 0..ngpu-1 (NCCL) */
  catt_ctx* ctx = NULL;
  catt_result* r = NULL;
  if (catt_ctx_create(&ctx, &cfg) != CATT_OK) { fprintf(stderr, "catt_ctx_create: %s\n", catt_last_error()); return 1; }
  if (catt_run_file(ctx, fastq, &r) != CATT_OK) { fprintf(stderr, "catt_run_file: %s\n", catt_last_error()); return 1; }
  catt_info info;
  catt_result_info(r, &info);
  unsigned long long h = 0xcbf29ce484222325ull;
  int64_t n_reads[2] = {0, 0}, with_cdr3 = 0;
  for (int which = 0; which < 2; which++) {
    const catt_read* rd = NULL;
    int64_t n = 0;
    if (catt_result_reads(r, which, &rd, &n) != CATT_OK) return 1;
    n_reads[which] = n;
    for (int64_t i = 0; i < n; i++) {
      const char* vn = rd[i].v_id >= 0 ? catt_ref_name(ctx, 0, rd[i].v_id) : "None";
      const char* jn = rd[i].j_id >= 0 ? catt_ref_name(ctx, 1, rd[i].j_id) : "None";
      if (strlen(rd[i].seq) != (size_t)rd[i].seq_len || strlen(rd[i].qual) != (size_t)rd[i].qual_len) { fprintf(stderr, "string length mismatch\n"); return 1; }
      h = fnv(h, rd[i].seq, (size_t)rd[i].seq_len); h = fnv(h, rd[i].qual, (size_t)rd[i].qual_len);
      h = fnv(h, vn, strlen(vn)); h = fnv(h, jn, strlen(jn));
      if (rd[i].cdr3_off >= 0) { h = fnv(h, rd[i].seq + rd[i].cdr3_off, (size_t)rd[i].cdr3_len); with_cdr3++; }
    }
  }
  /* the selBestMatchD loop (catt.jl:676-683) on the first CDR3 of Vpart */
  int32_t d = -2;
  {
    const catt_read* rd = NULL; int64_t n = 0;
    catt_result_reads(r, 0, &rd, &n);
    if (n > 0 && rd[0].cdr3_off >= 0) {
      char buf[1024];
      memcpy(buf, rd[0].seq + rd[0].cdr3_off, (size_t)rd[0].cdr3_len); buf[rd[0].cdr3_len] = 0;
      const char* one[1] = {buf};
      if (catt_best_d(ctx, one, 1, &d) != CATT_OK) return 1;
    }
  }
  /* the clonotype table: what nbsorb's grouping and the output counter need, without touching a read */
  const catt_clonotype* rows = NULL;
  int64_t n_rows = 0, row_reads = 0;
  if (catt_result_clonotypes(r, &rows, &n_rows) != CATT_OK) return 1;
  for (int64_t i = 0; i < n_rows; i++) {
    if (strlen(rows[i].cdr3) != (size_t)rows[i].cdr3_len || strlen(rows[i].qual) != (size_t)rows[i].cdr3_len) { fprintf(stderr, "clonotype row %lld malformed\n", (long long)i); return 1; }
    row_reads += rows[i].count;
  }
  if (row_reads != with_cdr3) { fprintf(stderr, "clonotype counts %lld != reads with a CDR3 %lld\n", (long long)row_reads, (long long)with_cdr3); return 1; }
  printf("clonotypes %lld ", (long long)n_rows);
  printf("reads %lld kmer %d the_err %.6e vpart %lld jpart %lld returned_v %lld returned_j %lld with_cdr3 %lld rescued %lld n_ranks %d best_d %d checksum %016llx\n",
         (long long)info.n_reads, info.kmer, info.the_err, (long long)info.vpart, (long long)info.jpart, (long long)n_reads[0], (long long)n_reads[1],
         (long long)with_cdr3, (long long)info.rescued, info.n_ranks, d, h);
  catt_result_free(r);
  catt_ctx_destroy(ctx);
  return 0;
}

int main(int argc, char** argv) {
  if (argc >= 2 && !strcmp(argv[1], "layout")) return layout();
  if (argc >= 4 && !strcmp(argv[1], "run")) return run(argv[2], argv[3], argc >= 5 ? atoi(argv[4]) : 1, argc >= 6 && !strcmp(argv[5], "distinct"));
  fprintf(stderr, "usage: roundtrip layout | roundtrip run <resource_dir> <fastq> [ngpu [distinct]]\n");
  return 2;
}
