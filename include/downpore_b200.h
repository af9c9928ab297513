/*
 * downpore_b200 — C ABI of the B200-native `downpore map` hot path.
 *
 * The reference (jteutenberg/downpore) has no FFI; the seam this library replaces is the Go interface
 *     mapping.Mapper { Map; MapWorker; AsString }          (mapping/mapping.go:22-26)
 * and its constructor
 *     mapping.NewMapper(reference, circular, k, kmerValues, seedRate, edgeSize, chunkSize, numWorkers)
 *                                                          (mapping/mapping.go:67)
 * as called from commands/map.go:73 and :84-86. A cgo shim binds exactly these symbols.
 *
 * Conventions
 *   - every function returns 0 on success, non-zero on failure; dp_last_error() returns a thread-local message
 *     (the reference only log.Fatal()s / panics on this path, so every error is fatal to the host);
 *   - input buffers are borrowed for the duration of the call only (cgo rule: no Go pointer is retained);
 *   - output buffers are allocated by the library and released with dp_free() — never free(): the large result arrays
 *     of the batch entry points are recycled through it (pages stay mapped from one call to the next);
 *   - a dp_mapper is bound to one CUDA device; create one per GPU and shard read batches across them
 *     (reads are independent: commands/map.go:84-86). The batch entry points may be called from several threads at once
 *     on one mapper, as the reference's num_workers goroutines call Mapper.Map (each call works on its own lanes — a
 *     stream and a workspace — of at most 12 per mapper; a caller that finds all of them taken waits);
 *     dp_mapper_get_stats reports the call that finished last;
 *   - there is no CPU fallback: every entry point fails if no CUDA device is usable.
 */
#ifndef DOWNPORE_B200_H
#define DOWNPORE_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct dp_mapper dp_mapper;

/* One mapping.Mapping (mapping/mapping.go:11-20) without the Query pointer (the caller knows which read it is). */
typedef struct dp_mapping {
    int64_t start;     /* Mapping.Start */
    int64_t end;       /* Mapping.End (inclusive coordinate for window hits: SURVEY Q3) */
    int32_t q_offset;  /* Mapping.QueryOffset */
    int32_t q_inset;   /* Mapping.QueryInset */
    int32_t ids;       /* Mapping.ids: reference bases covered by matched seeds */
    uint8_t rc;        /* Mapping.RC */
    uint8_t pad_[3];
} dp_mapping;

/* Device-side timings and work counters of the most recent dp_mapper_map_batch* call. */
typedef struct dp_stats {
    double ms_total;        /* wall time of the call inside the library */
    double ms_pack;         /* 2-bit pack kernel (only the queried windows are packed) */
    double ms_extract;      /* k-mer scan / seed extraction kernel, summed over rounds */
    double ms_lookup;       /* seed-index lookup (candidate chunks) kernel, summed over rounds */
    double ms_chain;        /* chaining kernels (sequential chaining, hand-backs, Map()'s first decision), summed over rounds */
    double ms_host_logic;   /* host wall time spent in the per-read Map() strategy between rounds */
    double ms_h2d;          /* host->device staging copies of pageable reads (0 for pinned or device-resident input) */
    int64_t rounds;         /* window-query rounds */
    int64_t windows;        /* performMapping window queries (both strands each) */
    int64_t kmer_lookups;   /* k-mer table gathers issued by the extract kernel */
    int64_t query_seeds;    /* seeds found in query window strands */
    int64_t posting_runs;   /* included seed occurrences (posting runs walked by the lookup kernel) */
    int64_t posting_entries;/* posting entries gathered by the lookup kernel */
    int64_t candidates;     /* candidate chunks produced by the lookup kernel */
    int64_t chain_cells;    /* |reduced chunk list| + |reduced query list| over candidates reaching the chainer */
    int64_t mappings;       /* mappings returned */
    int64_t kernel_launches;/* kernels launched by the call */
    int64_t bases;          /* sum of read lengths of the call */
    int64_t h2d_bytes;      /* read bytes that crossed the host->device link (copied, or - when the caller's buffer is
                               page-locked - only the queried windows, moved by the TMA pull kernel) */
    double ms_reduce;       /* list-reduction kernel of the chaining fast path (not included in ms_chain) */
    double ms_finish;       /* Map()'s first decision on the device (dp_finish_round0_kernel); included in ms_chain */
    int64_t retries;        /* sub-batch attempts recomputed with larger device capacities (repeat-rich references) */
    int64_t short_reads;    /* reads shorter than k+12 bases: returned without mappings (the reference's scans over-read
                               such slices: undefined there) */
} dp_stats;

/*
 * mapping.NewMapper (mapping/mapping.go:67-109): packs the reference, selects seed k-mers (AddSingleSeeds,
 * seeds/seeds.go:160-200), cuts the reference into chunks and builds the seed index, all on `device`.
 *   ref_ascii/ref_len : the first record of the reference FASTA (commands/map.go:34-36), one byte per base
 *   kmer_values       : 4^k doubles, commands/map.go:46-71 (stays an input so that Go's sort-dependent tie order at
 *                       the top-1 % boundary is inherited from the host, not re-implemented)
 *   seed_rate, edge_size (= query_size), chunk_size : as commands/map.go:39-43
 */
int dp_mapper_create(const uint8_t* ref_ascii, int64_t ref_len, int circular, int k, const double* kmer_values,
                     int seed_rate, int edge_size, int chunk_size, int device, dp_mapper** out);

/*
 * Mapper.Map over a batch of reads (mapping/mapping.go:430-487 for each read; replaces the MapWorker goroutine pool
 * of mapping.go:613-619 / commands/map.go:84-86).
 *   bases   : concatenated ASCII reads in host memory. Pinned / registered memory is read in place by the kernels
 *             (only the queried windows cross the link); pageable memory is staged through pinned buffers
 *   offsets : n_reads+1 byte offsets into `bases`
 *   out     : *out = malloc'ed array of all mappings, grouped per read in input order, each group in the slice order
 *             Map() returns them in
 *   out_offsets : *out_offsets = malloc'ed n_reads+1 prefix offsets into *out
 */
int dp_mapper_map_batch(dp_mapper* m, int64_t n_reads, const uint8_t* bases, const int64_t* offsets, dp_mapping** out,
                        int64_t** out_offsets);

/* Same, with `d_bases` (ASCII) already resident in device memory on the mapper's device; `offsets` stays a host array. */
int dp_mapper_map_batch_device(dp_mapper* m, int64_t n_reads, const uint8_t* d_bases, const int64_t* offsets,
                               dp_mapping** out, int64_t** out_offsets);

/*
 * Mapper.Map over a batch of reads that are PACKED ALREADY, in the reference's own in-memory form: the Go host's reader
 * hands the mapper `packedSequence`s (sequence/seqio.go:158,219 -> sequence.NewPackedSequence, sequence/sequence.go:67-93:
 * four bases per byte, first base in the two most significant bits, tail byte left-aligned and zero-padded), so this is
 * the entry a cgo host calls with the bytes it already holds and only a quarter of a byte per base crosses PCIe.
 * `packed` holds the reads' byte strings back to back or with gaps: read i starts at packed[byte_offsets[i]] and has
 * lengths[i] bases (byte_offsets has n_reads + 1 non-decreasing entries, the last one bounds the buffer). `packed` may be
 * pageable host memory, page-locked host memory (dp_host_alloc: read in place, only the queried windows cross the link)
 * or device memory. Results as dp_mapper_map_batch, identical to mapping the same reads as ASCII.
 */
int dp_mapper_map_batch_packed(dp_mapper* m, int64_t n_reads, const uint8_t* packed, const int64_t* byte_offsets,
                               const int64_t* lengths, dp_mapping** out, int64_t** out_offsets);

/* mapping.AsString (mapping/mapping.go:112-122): writes one PAF line (no newline) into buf; returns its length or -1. */
int dp_mapper_paf_line(const dp_mapper* m, const dp_mapping* mp, const char* query_name, int64_t query_len,
                       const char* ref_name, char* buf, int buf_len);

/*
 * Input and output side on the device.
 *
 * dp_split_records: the record rules of the reference's reader, readFasta's first pass (sequence/seqio.go:188-267),
 * applied on `device` to a FASTA/FASTQ file image — or to one piece of it — `image` in host or device memory:
 * the first line is a header; every later line that starts with 'A'..'T' is a sequence under the name line in force
 * (its line minus the last byte, kept when the line with its newline has at least min_length bytes); once a line
 * starting with '@' has been seen each sequence line is followed by a '+' line (anything else: error "Invalid fastq
 * format", the reference's log.Fatal) and a quality line. *records = malloc'ed table of name and sequence spans
 * (names TrimSpace'd, positions relative to `image`), in file order.
 * Pieces: with final_piece = 0 only the records in front of the piece's last name line are returned and *consumed is
 * where that line starts — hand the bytes from there on over again at the head of the next piece; *is_fastq (may be
 * NULL) carries the '@' state from piece to piece (start with 0). final_piece = 1: the image ends the file.
 */
typedef struct dp_record {
    int64_t name_start, name_len; /* the read's name (the reference's f.names entry) */
    int64_t seq_start, seq_len;   /* the read's bases */
} dp_record;
int dp_split_records(const uint8_t* image, int64_t bytes, int64_t min_length, int final_piece, int* is_fastq, int device,
                     dp_record** records, int64_t* n_records, int64_t* consumed);

/*
 * Mapper.Map over reads that are mapped where they lie in a file image (records of dp_split_records, ascending and
 * non-overlapping): no per-read copy on the host. `image` may be pageable or page-locked host memory or device memory.
 * Results as dp_mapper_map_batch on the same reads.
 */
int dp_mapper_map_batch_spans(dp_mapper* m, int64_t n_reads, const uint8_t* image, const dp_record* records,
                              dp_mapping** out, int64_t** out_offsets);

/*
 * Mapper.AsString (mapping/mapping.go:112-122) for every mapping of a batch, formatted on the device: *text = malloc'ed
 * block of PAF lines, each ended by '\n' (what commands/map.go:92 prints), in the order of `maps`; names are read from
 * `image` through `records`, the query length is the record's seq_len.
 */
int dp_mapper_paf_block(const dp_mapper* m, int64_t n_reads, const uint8_t* image, const dp_record* records,
                        const dp_mapping* maps, const int64_t* out_offsets, const char* ref_name, char** text,
                        int64_t* text_bytes);

/* Device memory for file images (a host uploads a piece once and hands the device pointer to the three calls above). */
int dp_device_alloc(void** out, size_t bytes, int device);
void dp_device_free(void* p, int device);
int dp_device_copy(void* dst, const void* src, size_t bytes, int device); /* any direction */

int dp_mapper_get_stats(const dp_mapper* m, dp_stats* out);

/* Index facts: out5 = {num_seeds, num_chunks, chunk_postings (seed occurrences over chunks),
 * seed_postings (distinct (seed,chunk) pairs), index_bytes on device}. */
int dp_mapper_index_info(const dp_mapper* m, int64_t* out5);

/*
 * Index image: the whole seed index (k-mer table, prefix filter, seed->chunk and seed->posting CSRs, chunk table) as
 * one relocatable byte image. It is what gets replicated when reads are sharded over several GPUs — build on one GPU,
 * export into a device buffer, broadcast it once over NVLink (NCCL), open a mapper from it on every other GPU — and
 * what an on-disk index is (export into host memory, write the bytes). `image` may be device or host memory in both
 * directions. A mapper opened from an image maps identically to the mapper it was exported from.
 */
int dp_mapper_index_image_size(const dp_mapper* m, int64_t* bytes);
int dp_mapper_index_export(const dp_mapper* m, void* image, int64_t bytes);
int dp_mapper_create_from_index(const void* image, int64_t bytes, int device, dp_mapper** out);

/* The mapper's parameters: out8 = {k, circular, ref_len, query_size, seed_rate, chunk_size, filter_bits, device}. */
int dp_mapper_params(const dp_mapper* m, int64_t* out8);

/* ------------------------------------------------------------------------------------------------------------------
 * `downpore overlap` (BASELINE config 5): one ROUND of commands/overlap.go:115-160 up to the stream of seed
 * matches overlapper.FindOverlaps delivers. The seam is the Go interface
 *     overlap.Overlapper { PrepareQueries; AddSequences; FindOverlaps; SetOverlapSize }   (overlap/overlap.go:24-29)
 * over a sequence.SequenceSet (sequence/seqio.go; himem: all reads cached) and a seeds.SeedIndex rebuilt every round
 * (commands/overlap.go:125-127). What stays on the host: the round loop, overlap.BuildConsensus over the matches of one
 * query and the PAF lines (commands/overlap.go:163-233), SequenceSet.SetIgnore.
 * Where the reference leaves an order to the goroutine scheduler (AddSeeds reads the k-mer table unlocked while other
 * workers write it; chunk ids and match order are arrival orders) the result is that of num_workers = 1: reads in file
 * order, chunks numbered in emission order, queries in slice order (forward, then reverse complement), candidates in
 * ascending chunk order.
 * ------------------------------------------------------------------------------------------------------------------ */
typedef struct dp_overlapper dp_overlapper;

/* one *seeds.SeedMatch of FindOverlaps' channel (overlap/overlap.go:374-376) */
typedef struct dp_overlap_hit {
    int32_t query_id; /* SeedMatch.QueryID: the slice (shared by its forward and reverse-complement query) */
    int32_t rc;       /* SeedMatch.ReverseComplementQuery */
    int32_t target;   /* chunk id: SeqB = index.GetSeedSequence(target) (dp_overlapper_chunks) */
    int32_t n;        /* len(MatchA) = len(MatchB) */
    int64_t at;       /* MatchA = matches[at .. at+n), MatchB = matches[at+n .. at+2n): seed indices in SeqA / SeqB */
} dp_overlap_hit;

typedef struct dp_overlap_round {
    int64_t num_seeds;           /* index.Size() after PrepareQueries */
    int64_t num_queries;         /* len(queries): two per slice; 0 ends the command (commands/overlap.go:132-134) */
    int64_t num_query_seqs;      /* numQuerySeqs (commands/overlap.go:136-145) */
    int64_t next_first_sequence; /* firstSequence of the next round */
    int64_t num_chunks;          /* index.GetNumSequences() after AddSequences */
    int64_t num_hits;
    int64_t num_matches;         /* entries of `matches` */
    dp_overlap_hit* hits;        /* malloc'ed, delivery order: queries ascending, targets ascending; dp_free */
    uint16_t* matches;           /* malloc'ed; dp_free */
    /* work counters and device timings (CUDA events on the overlapper's stream) */
    int64_t read_seeds;          /* seed occurrences over all reads sent to AddSequences */
    int64_t chunk_seeds;         /* seed occurrences over all chunks */
    int64_t seed_postings;       /* distinct (seed, chunk) pairs */
    int64_t posting_entries;     /* seed -> chunk postings gathered by the Matches kernel */
    int64_t candidates;          /* candidate chunks over all queries (SeedIndex.Matches) */
    int64_t pairs;               /* PairwiseAlignments calls (including the recomputed ones of a wave) */
    int64_t kernel_launches;
    double ms_total;             /* wall time of the call */
    double ms_select, ms_queries, ms_scan, ms_chunk, ms_index, ms_lookup, ms_align, ms_collect;
} dp_overlap_round;

/*
 * sequence.NewFastaSequenceSet (himem) + the parameters of commands/overlap.go:26-28: packs the reads onto `device`.
 *   bases_ascii / offsets : the reads that passed the min_length = overlap_size filter, concatenated (n_reads + 1 offsets)
 *   kmer_values           : 4^k doubles (getKmerValues, commands/overlap.go:41-95) or NULL: set later with
 *                           dp_overlapper_set_values after dp_overlapper_kmer_counts
 * FASTA only: FASTQ qualities (which weight AddSeeds' values, seeds/seeds.go:96-98) are not taken.
 */
int dp_overlapper_create(const uint8_t* bases_ascii, const int64_t* offsets, int64_t n_reads, int k,
                         const double* kmer_values, int overlap_size, int num_seeds, int seed_batch_size, int chunk_size,
                         int query_batch_size, double min_hits, int device, dp_overlapper** out);
/* sequtil.KmerOccurrences over every read (commands/overlap.go:43): counts[4^k] += occurrences */
int dp_overlapper_kmer_counts(dp_overlapper* o, uint64_t* counts);
int dp_overlapper_set_values(dp_overlapper* o, const double* kmer_values);
/*
 * One round: PrepareQueries(num_seeds, seed_batch_size, values, GetNSequencesFrom(first_sequence, query_batch_size),
 * QueryEdges), AddSequences(GetSequences()), FindOverlaps(queries).  ignore: n_reads flags (SequenceSet.SetIgnore) or NULL.
 */
int dp_overlapper_round(dp_overlapper* o, const uint8_t* ignore, int64_t first_sequence, dp_overlap_round* out);
/* The seed sequences of the last round, as the reference's segments (gap, seed, gap, ..., gap) with its seed ids:
 *   queries: meta = num_queries x {ID, SequenceID, rc, length, offset, inset}; seg_off[num_queries + 1]; *segs malloc'ed
 *   chunks : ids (NULL = all num_chunks); meta = n x {read, length, offset, inset, n_seeds}; seg_off[n + 1]; *segs malloc'ed
 *   seed_kmers: the k-mer of every seed id (seedMap), num_seeds entries */
int dp_overlapper_queries(dp_overlapper* o, int64_t* meta, int64_t* seg_off, int64_t** segs);
int dp_overlapper_chunks(dp_overlapper* o, const int32_t* ids, int64_t n, int64_t* meta, int64_t* seg_off, int64_t** segs);
int dp_overlapper_seed_kmers(dp_overlapper* o, int64_t* kmers_out);
void dp_overlapper_destroy(dp_overlapper* o);

/* Test probes (stage dumps for parity tests; not needed by a host). */
/* seed k-mers, ascending k-mer value; out must hold num_seeds entries */
int dp_mapper_seed_kmers(const dp_mapper* m, int64_t* out);
/* chunk c: fields4 = {offset, inset, length, n_seeds}; pos/kmer (n_seeds each, may be NULL) */
int dp_mapper_chunk(const dp_mapper* m, int64_t c, int64_t* fields4, int32_t* pos, int64_t* kmer);
/*
 * One window query (performMapping, mapping/mapping.go:489-611) on read bases[0..read_len): window [start,end) or the
 * whole read. Any of the outputs may be NULL. seeds: per strand (0 fwd, 1 rc) n, then (pos, kmer) pairs;
 * candidates per strand; mappings as returned by performMapping.
 */
int dp_mapper_probe_window(dp_mapper* m, const uint8_t* read_ascii, int64_t read_len, int64_t start, int64_t end,
                           int whole, int32_t* n_seeds2, int32_t* seed_pos, int64_t* seed_kmer, int64_t seed_cap,
                           int32_t* n_cand2, int32_t* cand, int64_t cand_cap, int32_t* n_map, dp_mapping* maps,
                           int64_t map_cap);

/* 2-bit packing of one sequence exactly as sequence.NewPackedSequence (sequence/sequence.go:67-93) lays it out:
 * 4 bases per byte, MSB first, tail byte left-aligned; out must hold (len+3)/4 bytes. Runs the device pack kernel. */
int dp_pack(const uint8_t* ascii, int64_t len, uint8_t* out, int device);

/* sequtil.KmerOccurrences (util/sequtil/kmers.go:34-69) for one record: counts[4^k] += occurrences. */
int dp_kmer_counts(const uint8_t* ascii, int64_t len, int k, uint64_t* counts, int device);

/*
 * Measurement aid (bench.py): random 32-byte-sector gather bandwidth over a `table_bytes` table (choose it far larger
 * than L2) in GB/s of sector traffic: the HBM gather roofline the index-lookup kernel is held against.
 */
int dp_probe_gather_gbs(int device, int64_t table_bytes, double* sector_gbs);

/*
 * Page-locked host memory for read batches (what a cgo host passes as `bases`): dp_mapper_map_batch reads such a
 * buffer in place from the device (only the queried windows cross the link: TMA bulk copies out of the mapped buffer),
 * so a host that fills batches into dp_host_alloc'ed memory never pays a staging copy. Portable across devices.
 * Release with dp_host_free.
 */
int dp_host_alloc(void** out, size_t bytes);
void dp_host_free(void* p);

void dp_mapper_destroy(dp_mapper* m);
void dp_free(void* p);
const char* dp_last_error(void);
/* library version string, e.g. "downpore_b200 0.1 (sm_100a)" */
const char* dp_version(void);

#ifdef __cplusplus
}
#endif
#endif /* DOWNPORE_B200_H */
