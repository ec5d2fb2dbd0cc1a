/*
 * mirge_b200.h -- C ABI of the B200-native miRge2.0 annotate-mode hot path.
 *
 * The reference (mhalushka/miRge, Python 2) has no FFI: its hot path is reached
 * through Python function boundaries and through process/file boundaries to
 * cutadapt and bowtie.  Each entry point below names the reference interface it
 * replaces (paths relative to /root/reference/src/mirge).  Plain pointers and
 * sizes only; one context per process per GPU; not thread-safe.  Every function
 * returns 0 on success or a negative MIRGE_E_* code, with a message available
 * from mirge_last_error().  Caller-owned buffers are never freed or retained.
 */
#ifndef MIRGE_B200_H
#define MIRGE_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct mirge_ctx mirge_ctx;

enum {
    MIRGE_OK = 0,
    MIRGE_E_ARG = -1,       /* bad argument */
    MIRGE_E_CUDA = -2,      /* CUDA runtime error */
    MIRGE_E_FORMAT = -3,    /* FASTQ format error (cutadapt FormatError) */
    MIRGE_E_TOOLONG = -4,   /* a read exceeds max_read_len */
    MIRGE_E_FULL = -5,      /* unique-sequence table full (raise max_unique) */
    MIRGE_E_STATE = -6      /* call order violated */
};

/* Library slots; rounds use them in the order of
 * utils/runAnnotationPipeline.py:588-599 (spike-in variant :574-586). */
enum {
    MIRGE_LIB_MIRNA = 0, MIRGE_LIB_HAIRPIN = 1, MIRGE_LIB_MATURE_TRNA = 2, MIRGE_LIB_PRE_TRNA = 3,
    MIRGE_LIB_SNORNA = 4, MIRGE_LIB_RRNA = 5, MIRGE_LIB_NCRNA_OTHERS = 6, MIRGE_LIB_MRNA = 7,
    MIRGE_LIB_GENOME = 8, MIRGE_LIB_SPIKEIN = 9, MIRGE_N_LIBS = 10
};

#define MIRGE_N_ROUNDS 10       /* 9 rounds, 10 with -spikeIn */
#define MIRGE_N_TOTALS 12       /* columns of lib_totals, see mirge_export_tables */

/* -- context ------------------------------------------------------------- */

/* Replaces the four dicts main() owns (__main__.py:318-330: seqDic, mirDic,
 * readLengthDic, logDic) with device-resident state on CUDA device `device`.
 * max_unique bounds the number of distinct trimmed sequences; max_read_len
 * (<= 128) bounds the raw read length. */
int mirge_create(int device, int n_samples, int spike_in, uint64_t max_unique, int max_read_len, mirge_ctx **out);
void mirge_destroy(mirge_ctx *ctx);
const char *mirge_last_error(const mirge_ctx *ctx);

/* Run all work on an existing cudaStream_t (passed as void*); default is a
 * private non-blocking stream. */
int mirge_set_stream(mirge_ctx *ctx, void *cuda_stream);
/* Wait for queued work and surface device-side error flags. */
int mirge_sync(mirge_ctx *ctx);

/* Return to the freshly created state (libraries, adapter and stream are
 * kept): what starting a new `miRge2.0 annotate` run does to the four dicts. */
int mirge_reset(mirge_ctx *ctx);

/* -- stage 1: trim + quality/length filter + collapse ------------------- */

/* Worker.__init__ (utils/trim_file.py:24-41): spec is the `adapter` string
 * trim_file receives -- "none", "+N" (UnconditionalCutter) or one or more
 * comma-separated 3' adapters (ACGT only, each <= 64 nt);
 * error_rate 0.12 (:32), min_overlap 3 (cutadapt default), min_length 16 (:33),
 * qual_cutoff 10 (:30). */
int mirge_set_adapter(mirge_ctx *ctx, const char *spec, double error_rate, int min_overlap, int min_length, int qual_cutoff);

/* Quality base auto-detection as it takes effect in trim_file (:104-109):
 * 64 iff the FIRST read holds a quality character > 'J'.  Host-side, pure. */
int mirge_detect_phred(const char *fastq, uint64_t nbytes);

/* trim_file(infile, adapter, outfile, threads, phred) (utils/trim_file.py:89)
 * fused with quantReads(...) (utils/quantReads.py:3-24) for one record-aligned
 * chunk of FASTQ text (4-line records) of sample `sample`:  H2D copy, parse and
 * 2-bit pack (K0), then 3' quality trim, 3' adapter removal, length filter and
 * insertion into the device hash table of unique sequences (K1).  No trimmed
 * FASTQ is written.  The text is cut into 64 MiB regions whose H2D copies are
 * all queued up front, so transfers overlap the kernels of earlier regions
 * (pinned memory for true overlap).  All copies have completed when the call
 * returns; the last region's kernels may still be running (mirge_sync). */
int mirge_feed_fastq(mirge_ctx *ctx, int sample, const char *fastq, uint64_t nbytes, int phred_base);
/* Same, FASTQ text already resident in device memory (16-byte aligned). */
int mirge_feed_fastq_device(mirge_ctx *ctx, int sample, const void *fastq_dev, uint64_t nbytes, int phred_base);

/* (processed, kept) per sample as returned by trim_file (:134) and the number
 * of unique sequences (len(seqDic)).  Synchronises. */
int mirge_get_stats(mirge_ctx *ctx, uint64_t *processed, uint64_t *kept, uint64_t *n_unique);

/* -- stage 2: mapping rounds --------------------------------------------- */

/* Stands in for the bowtie index files main() resolves (__main__.py:262-281):
 * n_refs entries, sequences concatenated as ASCII with offsets[n_refs+1].
 * Names stay with the caller; kernels speak reference indices.  Builds the
 * 2-bit reference and its k-mer seed tables on the device. */
int mirge_load_library(mirge_ctx *ctx, int lib_id, int n_refs, const char *seq_ascii, const uint64_t *offsets);

/* runAnnotationPipeline(...) (utils/runAnnotationPipeline.py:566,636-705): all
 * rounds in order over the unique sequences that have not been looked at yet
 * (incremental: sequences added by later feeds are picked up by the next call;
 * loading a library or rebuilding the table starts over), including the poly-T
 * pre-tRNA round (:664-704). */
int mirge_annotate(mirge_ctx *ctx);

/* -- genome passes --------------------------------------- */

/* index.Libs/<sp>_genome.*.ebwt (__main__.py:270): chromosomes concatenated as
 * ASCII with offsets[n_refs+1]; builds the 2-bit genome and ONE prefix-ordered
 * k-mer table that serves every seed length (csrc/k2_genome.cuh). */
int mirge_load_genome(mirge_ctx *ctx, int n_refs, const char *seq_ascii, const uint64_t *offsets);
/* `bowtie <genome> -n N|-v V -f -a [-3 trim3] [-l seedlen] [--norc]` for n reads
 * (ASCII, `stride` bytes apart, lens[i] bases each; host memory):
 *   utils/writeDataToCSV.py:1260  mode 0, param 1, seedlen 28, trim3 2   (A-to-I candidates)
 *   utils/writeDataToCSV.py:1488  mode 0, param 0, seedlen 28, trim3 2
 *   __main__.py:538               mode 0, param 0, seedlen 25, trim3 0   (-m 3 --best: caller suppresses sum > 3)
 * counts[i*4 + k] (host) = number of alignments of read i, both strands unless
 * norc, with k mismatches.  What bowtie would print per read follows from the
 * counts (-a: all of them; --best --strata: the lowest occupied k; -m M:
 * nothing when the sum exceeds M).  Synchronises. */
int mirge_genome_pass(mirge_ctx *ctx, uint64_t n, const char *reads, int stride, const uint8_t *lens,
                      int mode, int param, int seedlen, int trim3, int norc, uint32_t *counts);
/* Same for unique sequences already in the device table, named by their ids
 * (row numbers of mirge_export_unique). */
int mirge_genome_pass_ids(mirge_ctx *ctx, uint64_t n, const uint32_t *ids, int mode, int param, int seedlen, int trim3, int norc,
                          uint32_t *counts);

/* Options.  MIRGE_OPT_ANNOTATE_WHILE_FEEDING (default off): mirge_feed_fastq runs the
 * mapping rounds on the sequences each 64 MiB region added while the next region
 * is being copied, so mirge_annotate only has the last region left -- the rounds
 * hide behind the host-to-device copies.  Libraries must be loaded before
 * feeding; leave it off when the table is rebuilt afterwards (multi-GPU exchange).
 * MIRGE_OPT_PAIR_SEEDS (default 16; read by mirge_load_library): short queries of
 * the -n 1 / -v 1 / -v 2 rounds are searched through pair-of-pieces seed tables
 * wherever the plain pigeonhole seeds would offer at least this many candidate
 * placements per query; 1 = wherever such a table can serve, 0 = never.  The
 * tables only shorten the candidate lists -- the alignments reported are the same. */
enum { MIRGE_OPT_ANNOTATE_WHILE_FEEDING = 1, MIRGE_OPT_PAIR_SEEDS = 2 };
int mirge_set_option(mirge_ctx *ctx, int option, int value);
/* Which pair-of-pieces seed sets mirge_load_library builds for a library of `total`
 * bases in slot lib_id under MIRGE_OPT_PAIR_SEEDS = pair_seeds: up to four rows of
 * {pieces, piece length, shortest region served, longest region served} in out16
 * (may be NULL); returns the number of rows.  Host arithmetic only, no device call. */
int mirge_pair_seed_plan(int lib_id, uint64_t total, int pair_seeds, int *out16);

/* -- stage 3: tallies ----------------------------------------------------- */

/* summarize(...) integer part (utils/summarize.py:34-66). */
int mirge_tally(mirge_ctx *ctx);

/* -- export ---------------------------------------------------------------- */

/* seqDic (__main__.py:318-321): for unique u < n_unique, seq[u*stride ..] holds
 * the ASCII sequence (no terminator), len[u] its length, counts[u*S+s] the
 * per-sample read count, round_of[u] = 0 (un-annotated) or round index + 1 as in
 * annot[] (runAnnotationPipeline.py:345), ref_of/pos_of the reference entry and
 * 0-based offset of the reported hit (for round 9/isomiR: of the -5 1 -3 2
 * trimmed read), mm_of its mismatch count.  Any pointer may be NULL. */
int mirge_export_unique(mirge_ctx *ctx, uint64_t cap, int stride, char *seq, uint8_t *len, uint32_t *counts,
                        int8_t *round_of, int32_t *ref_of, int32_t *pos_of, int8_t *mm_of);

/* mir_quant / mir_iscan: [n_mirna_refs][S] (mirDic 'quant' / 'iscan' before
 * merge+filter).  lib_totals: [S][12] = trimmedUniq, mirnaReads, hairpinReads,
 * maturetrnaReads, pretrnaReads, snornaReads, rrnaReads, ncrnaOthersReads,
 * mrnaReads, spikeInReads, remReads, (unused)  (logDic['quantStats']).
 * len_hist: [max_read_len+1][S] (readLengthDic, quantReads.py:17-21).
 * round_processed / round_aligned: [10] (parseBowtieLog, :9-18). */
int mirge_export_tables(mirge_ctx *ctx, int64_t *mir_quant, int64_t *mir_iscan, int64_t *lib_totals,
                        uint64_t *len_hist, uint64_t *round_processed, uint64_t *round_aligned);

/* -- multi-GPU exchange (no reference counterpart; samples are serial there) */

/* Bytes of one exchange record: packed key, then len | sample << 8 and the read
 * count of that sample (32 bytes for reads <= 64 nt). */
int mirge_record_bytes(const mirge_ctx *ctx);
/* X1, sender side: write every (unique sequence, sample) pair with
 * a non-zero count as a record into dev_buf (capacity cap_records), grouped by
 * owner rank = hash(key) % n_parts; part_counts_dev[n_parts] (DEVICE memory)
 * receives the group sizes.  Asynchronous on the context's stream: no host
 * round trip, the caller moves the sizes and the records with its transport
 * (mirge_b200/dist.py: torch.distributed all-to-all over NCCL). */
int mirge_export_records_device(mirge_ctx *ctx, int n_parts, void *dev_buf, uint64_t cap_records, uint64_t *part_counts_dev);
/* Forget all unique sequences (per-sample processed/kept and len_hist stay). */
int mirge_reset_unique(mirge_ctx *ctx);
/* X1, owner side: insert n_records received records, adding their counts. */
int mirge_import_records_device(mirge_ctx *ctx, const void *dev_buf, uint64_t n_records);
/* X2: the integer tables of this rank -- mir_quant, mir_iscan,
 * lib_totals, len_hist, kept, processed, in this order -- as one contiguous
 * int64 vector of mirge_tables_words() words in device memory, for an in-place
 * sum over ranks (NCCL all-reduce), and back.  Asynchronous on the stream. */
uint64_t mirge_tables_words(const mirge_ctx *ctx);
int mirge_pack_tables_device(mirge_ctx *ctx, void *dst_dev);
int mirge_unpack_tables_device(mirge_ctx *ctx, const void *src_dev);
/* Owner hash of an ASCII sequence (for host-side tests of the partitioning). */
uint64_t mirge_hash_sequence(const char *seq, int len);

/* -- measurement ----------------------------------------------------------- */

/* Per-kernel CUDA-event timing on the context's stream. */
#define MIRGE_N_KERNELS 32
int mirge_profile_enable(mirge_ctx *ctx, int on);
/* ms[k], launches[k] accumulated since enable; names[k] static strings. */
int mirge_profile_read(mirge_ctx *ctx, double *ms, uint64_t *launches, const char **names);
/* Test hook: force_slow = 1 sends every read through the slow queue (all-column
 * bit-vector search), 3 through the literal banded cutadapt DP, 4 switches the
 * exact-match prefilter off (every read takes the windowed bit-vector search),
 * 0x2000 parses FASTQ with the exact chained-scan kernel only (no speculation),
 * 0x1000 routes every miss through the poly-T round (no bypass list),
 * 0 restores the default, -1 leaves the setting unchanged; slow_reads (may be
 * NULL) receives the number of reads that went through the slow queue so far.
 * Results are identical in every mode. */
int mirge_debug(mirge_ctx *ctx, int force_slow, uint64_t *slow_reads);
/* Diagnostics since the last reset: out4 = reads through the all-column search,
 * reads of prefilter class 1 (settled or passed on by the piece filter), reads sent
 * to the general search (class 2), feed regions whose speculative parse was
 * rejected and redone by the exact kernel (k0_parse.cuh).  Synchronises. */
int mirge_trim_classes(mirge_ctx *ctx, uint64_t *out4);
/* The tables behind the trim kernel's exact-match prefilter for ONE 3' adapter
 * (mirge_b200/csrc/k1_prefilter.cuh): returns 1 and fills m, icap, safe_p,
 * safe_len[49] (index = insert length I, bit l) and fullreach[32] (index = l)
 * when the prefilter applies (8 <= adapter length <= 32), 0 when every read takes
 * the general search, negative on a bad argument.  Host-side, pure; lets the CPU
 * tests check the tables against the oracle without a GPU. */
int mirge_prefilter_tables(const char *adapter, double error_rate, int min_overlap, int *m, int *icap,
                           uint32_t *safe_p, uint32_t *safe_len, uint8_t *fullreach);
/* Number of kernels this library has launched on the context so far. */
uint64_t mirge_launch_count(const mirge_ctx *ctx);

#ifdef __cplusplus
}
#endif
#endif /* MIRGE_B200_H */
