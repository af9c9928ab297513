// dp_map — C++ host of `downpore map` over the downpore_b200 C ABI (include/downpore_b200.h).
//
// Stands where the reference's Go host stands (commands/map.go:33-116, downpore.go:34-51, sequence/seqio.go:188-267):
// same arguments, aliases and defaults, same record parsing rules, same PAF records and stderr counters. The Go
// toolchain is absent from the build image, so this is the host the tests and timings drive; a Go maintainer would
// bind the same symbols through cgo. All mapping work happens on the GPU behind the ABI: this file
// only reads files, batches reads and prints.
//
//   dp_map -input reads.fasta -reference ref.fasta [-circular true] [-k 11] [-query_size 1000] [-min_length 500]
//          [-chunk_size 10000] [-seed_rate 40] [-num_workers 4]
//
// Environment (the flag surface stays the reference's): DOWNPORE_GPUS = comma separated device ordinals (default 0),
// DOWNPORE_BATCH / DOWNPORE_BATCH_BYTES = reads / bytes per dp_mapper_map_batch call (default 131072 / 1.5 GiB), DOWNPORE_STATS=1 prints stage timings.
// Output order: records grouped per read in input order (the reference prints in goroutine completion order).
#include <algorithm>
#include <array>
#include <chrono>
#include <condition_variable>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <numeric>
#include <string>
#include <thread>
#include <vector>

#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include "../include/downpore_b200.h"

namespace {

[[noreturn]] void fatal(const std::string& msg) {
    fprintf(stderr, "%s\n", msg.c_str());
    exit(1);
}

// ---- file access --------------------------------------------------------------------------------------------------
struct MappedFile {
    const unsigned char* p = nullptr;
    size_t n = 0;
    int fd = -1;
    explicit MappedFile(const std::string& path) {
        fd = open(path.c_str(), O_RDONLY);
        if (fd < 0) fatal("open " + path + ": " + strerror(errno));
        struct stat st;
        if (fstat(fd, &st) != 0) fatal("stat " + path);
        n = (size_t)st.st_size;
        if (n) {
            void* m = mmap(nullptr, n, PROT_READ, MAP_PRIVATE, fd, 0);
            if (m == MAP_FAILED) fatal("mmap " + path);
            madvise(m, n, MADV_SEQUENTIAL);
            p = (const unsigned char*)m;
        }
    }
    ~MappedFile() {
        if (p) munmap((void*)p, n);
        if (fd >= 0) close(fd);
    }
};

// ---- record reader: the first-pass rules of readFasta (sequence/seqio.go:188-267) ----------------------------------
// One line per sequence. The first line is a header; a later line is a sequence iff its first byte is in 'A'..'T';
// a record is kept iff its line length including the newline is >= min_length; the sequence is the line minus its last
// byte (also when the final line has no newline); the name is the previous header without its first byte, trimmed.
// A leading '@' switches to FASTQ: after each sequence line a '+' line and a quality line follow.
struct Record {
    const unsigned char* seq;
    size_t len;
    const unsigned char* name;
    size_t nameLen;
};

class RecordReader {
public:
    RecordReader(const unsigned char* p, size_t n, long long minLength) : p_(p), n_(n), minLength_(minLength) {
        size_t len;
        bool eof;
        const unsigned char* line = next_line(len, eof);
        if (!line || eof) {
            done_ = true;
            return;
        }
        if (line[0] == '@') fastq_ = true;
        set_name(line, len);
    }
    bool next(Record& r) {
        while (!done_) {
            size_t len;
            bool eof;
            const unsigned char* line = next_line(len, eof);
            if (!line || len == 0) {
                done_ = true;
                return false;
            }
            bool have = false;
            if (line[0] >= 'A' && line[0] <= 'T') {
                if ((long long)len >= minLength_) {
                    r.seq = line;
                    r.len = len - 1;
                    r.name = name_;
                    r.nameLen = nameLen_;
                    have = true;
                }
                if (fastq_) {
                    size_t l2;
                    bool e2;
                    const unsigned char* plus = next_line(l2, e2);
                    if (!plus || e2 || plus[0] != '+') fatal("Invalid fastq format (on + line)");
                    next_line(l2, e2);  // quality line; the loop's err is the '+' line's (nil) from here on
                    eof = false;
                }
            } else if (line[0] == '@') {
                fastq_ = true;
                set_name(line, len);
            } else {
                set_name(line, len);
            }
            if (eof) done_ = true;
            if (have) return true;
        }
        return false;
    }

private:
    const unsigned char* next_line(size_t& len, bool& eof) {  // bufio.ReadBytes('\n'): the line including its '\n'
        if (pos_ >= n_) {
            len = 0;
            eof = true;
            return nullptr;
        }
        const unsigned char* s = p_ + pos_;
        const void* nl = memchr(s, '\n', n_ - pos_);
        if (nl) {
            len = (size_t)((const unsigned char*)nl - s) + 1;
            eof = false;
        } else {
            len = n_ - pos_;
            eof = true;
        }
        pos_ += len;
        return s;
    }
    void set_name(const unsigned char* line, size_t len) {  // strings.TrimSpace(string(line[1:]))
        const unsigned char* a = line + 1;
        const unsigned char* b = line + len;
        auto sp = [](unsigned char c) { return c == ' ' || (c >= '\t' && c <= '\r'); };
        while (a < b && sp(*a)) a++;
        while (b > a && sp(b[-1])) b--;
        name_ = a;
        nameLen_ = (size_t)(b - a);
    }
    const unsigned char* p_;
    size_t n_;
    long long minLength_;
    size_t pos_ = 0;
    bool fastq_ = false, done_ = false;
    const unsigned char* name_ = nullptr;
    size_t nameLen_ = 0;
};

// ---- values[] of commands/map.go:46-71 ----------------------------------------------------------------------------
uint64_t revcomp_id(uint64_t x, int k) {
    uint64_t r = 0;
    for (int i = 0; i < k; i++) {
        r = (r << 2) | ((x ^ 3) & 3);
        x >>= 2;
    }
    return r;
}

std::vector<double> kmer_values(std::vector<uint64_t>& counts, int k) {
    const size_t n = counts.size();
    std::vector<double> values(n);
    uint64_t tot = 0;
    for (uint64_t c : counts) tot += c;
    const double tf = (double)tot, target = 0.000005;
    for (size_t i = 0; i < n; i++) {
        double freq = (double)counts[i] / tf;
        if (counts[i] < 3) values[i] = 0;
        else if (freq <= target) values[i] = 1.0 - (target - freq);
        else values[i] = 1.0 - (freq - target);
    }
    // TopOccurrences(counts, k, n/100, n/50) (util/sequtil/kmers.go:87-112): forward and reverse-complement counts are
    // merged in place over ascending ids, the ids sorted by merged count, the top n/100 zeroed. Ties at the cut follow
    // Go's sort.Sort in the reference (unpinned); here they are broken by ascending id.
    for (size_t i = 0; i < n; i++) {
        uint64_t rc = revcomp_id(i, k);
        uint64_t c = counts[i] + counts[rc];
        counts[i] = c;
        counts[rc] = c;
    }
    // (only the SET of the n/100 largest under the total order (count, id) is needed: a selection, not a sort)
    std::vector<uint32_t> ids(n);
    std::iota(ids.begin(), ids.end(), 0u);
    auto less = [&](uint32_t a, uint32_t b) { return counts[a] != counts[b] ? counts[a] < counts[b] : a < b; };
    std::nth_element(ids.begin(), ids.begin() + (n - n / 100), ids.end(), less);
    for (size_t i = n - n / 100; i < n; i++) values[ids[i]] = 0;
    values[0] = 0;
    return values;
}

// ---- batches ------------------------------------------------------------------------------------------------------
struct Batch {
    uint8_t* bases = nullptr;  // pinned (dp_host_alloc): the kernels pull the queried windows straight out of it
    size_t cap = 0, used = 0;
    std::vector<int64_t> offsets;
    std::vector<std::pair<const unsigned char*, size_t>> names;
    std::vector<const unsigned char*> srcs;  // where each read's bases lie in the input file
    bool last = false;
    void reset() {
        used = 0;
        offsets.assign(1, 0);
        names.clear();
        srcs.clear();
        last = false;
    }
};

// fn(t, lo, hi) over [0, n) cut into contiguous pieces, one per thread
template <class F>
void parallel_ranges(size_t n, size_t threads, F fn) {
    threads = std::max<size_t>(1, std::min(threads, n));
    if (threads == 1) {
        fn((size_t)0, (size_t)0, n);
        return;
    }
    std::vector<std::thread> th;
    for (size_t t = 0; t < threads; t++) th.emplace_back([=] { fn(t, n * t / threads, n * (t + 1) / threads); });
    for (auto& x : th) x.join();
}

double now_s() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

}  // namespace

int main(int argc, char** argv) {
    // downpore.go:34-51 parseArgs + commands/command.go:18-58 MakeArgs (shortest unambiguous prefixes as aliases)
    std::map<std::string, std::string> args = {{"input", ""},          {"reference", ""},      {"circular", "true"},
                                               {"k", "11"},            {"query_size", "1000"}, {"min_length", "500"},
                                               {"chunk_size", "10000"}, {"seed_rate", "40"},   {"num_workers", "4"}};
    const std::map<std::string, std::string> alias = {{"i", "input"},       {"r", "reference"},  {"ci", "circular"},
                                                      {"k", "k"},           {"q", "query_size"}, {"m", "min_length"},
                                                      {"ch", "chunk_size"}, {"s", "seed_rate"},  {"n", "num_workers"}};
    int first = 1;
    if (argc > 1 && strcmp(argv[1], "map") == 0) first = 2;  // accept `dp_map map -input ...` like `downpore map ...`
    for (int i = first; i < argc; i += 2) {
        std::string name = argv[i];
        name.erase(0, name.find_first_not_of('-'));
        auto a = alias.find(name);
        if (a != alias.end()) name = a->second;
        if (!args.count(name)) fatal("Unrecognised argument:" + name);
        if (i + 1 >= argc) fatal("Missing value for argument:" + name);
        args[name] = argv[i + 1];
    }
    auto parse_int = [](const std::string& s) {  // commands.ParseInt: base 10, 32 bits
        char* end = nullptr;
        errno = 0;
        long long v = strtoll(s.c_str(), &end, 10);
        if (s.empty() || *end || errno || v > INT32_MAX || v < INT32_MIN) fatal("Invalid integer argument value:" + s);
        return (int)v;
    };
    const int k = parse_int(args["k"]);
    parse_int(args["num_workers"]);  // validated like the reference; the goroutine pool is replaced by GPU batches
    const int minLength = parse_int(args["min_length"]);
    const std::string& cs = args["circular"];
    const bool circular = cs == "1" || (!cs.empty() && (cs[0] == 'T' || cs[0] == 't'));
    const int querySize = parse_int(args["query_size"]);
    const int chunkSize = parse_int(args["chunk_size"]);
    const int seedRate = parse_int(args["seed_rate"]);
    if (args["input"].empty() || args["reference"].empty()) fatal("usage: dp_map -input <fasta/fastq> -reference <fasta> [...]");
    if (k < 1 || k > 15) fatal("k must be in [1, 15]");

    std::vector<int> devices;
    {
        const char* env = getenv("DOWNPORE_GPUS");
        std::string s = env ? env : "0";
        size_t pos = 0;
        while (pos <= s.size()) {
            size_t c = s.find(',', pos);
            if (c == std::string::npos) c = s.size();
            if (c > pos) devices.push_back(atoi(s.substr(pos, c - pos).c_str()));
            pos = c + 1;
        }
        if (devices.empty()) devices.push_back(0);
    }
    const size_t batchReads = getenv("DOWNPORE_BATCH") ? (size_t)atoll(getenv("DOWNPORE_BATCH")) : (size_t)131072;
    const size_t batchBytes =
        getenv("DOWNPORE_BATCH_BYTES") ? (size_t)atoll(getenv("DOWNPORE_BATCH_BYTES")) : ((size_t)3 << 29);
    const bool wantStats = getenv("DOWNPORE_STATS") != nullptr;
    // host threads for copying reads into the batch buffers and for formatting PAF lines (DOWNPORE_HOST_THREADS)
    const size_t hostThreads = getenv("DOWNPORE_HOST_THREADS")
                                   ? (size_t)std::max(1, atoi(getenv("DOWNPORE_HOST_THREADS")))
                                   : (size_t)std::max(1u, std::min(8u, std::thread::hardware_concurrency() / 2));

    // ---- reference: the first record is indexed, every record is counted (commands/map.go:34-36, 45) ----
    double t0 = now_s();
    MappedFile refFile(args["reference"]);
    std::string refName;
    std::vector<uint8_t> reference;
    std::vector<uint64_t> counts((size_t)1 << (2 * k), 0);
    {
        RecordReader rr(refFile.p, refFile.n, 0);
        Record r;
        bool firstRec = true;
        while (rr.next(r)) {
            if (firstRec) {
                reference.assign(r.seq, r.seq + r.len);
                refName.assign((const char*)r.name, r.nameLen);
                firstRec = false;
            }
            if ((long long)r.len >= k && dp_kmer_counts(r.seq, (int64_t)r.len, k, counts.data(), devices[0]))
                fatal(std::string("dp_kmer_counts: ") + dp_last_error());
        }
        if (firstRec) fatal("no reference sequence");
    }
    std::vector<double> values = kmer_values(counts, k);
    fprintf(stderr, "K-mer counting complete. Preparing to start indexing and querying...\n");
    double t1 = now_s();

    std::vector<dp_mapper*> mappers(devices.size(), nullptr);
    {   // one replica of the index per GPU, built concurrently (deterministic, so every replica is identical)
        std::vector<std::thread> th;
        std::vector<std::string> errs(devices.size());
        for (size_t d = 0; d < devices.size(); d++)
            th.emplace_back([&, d] {
                if (dp_mapper_create(reference.data(), (int64_t)reference.size(), circular ? 1 : 0, k, values.data(),
                                     seedRate, querySize, chunkSize, devices[d], &mappers[d]))
                    errs[d] = dp_last_error();
            });
        for (auto& t : th) t.join();
        for (auto& e : errs)
            if (!e.empty()) fatal("dp_mapper_create: " + e);
    }
    double t2 = now_s();

    MappedFile in(args["input"]);
    if (getenv("DOWNPORE_DEVICE_IO") && atoi(getenv("DOWNPORE_DEVICE_IO")) != 0) {
        // ---- the host-free route: the file goes to the GPUs piece by piece as it is; records are split
        // (dp_split_records), reads mapped where they lie (dp_mapper_map_batch_spans) and PAF lines formatted
        // (dp_mapper_paf_block) on the device; the host only cuts pieces, counts and writes the text in piece order ----
        const size_t pieceBytes = getenv("DOWNPORE_PIECE_BYTES") ? (size_t)atoll(getenv("DOWNPORE_PIECE_BYTES")) : ((size_t)1 << 30);
        std::mutex mu;
        std::condition_variable cv;
        size_t filePos = 0;
        int fq = 0;
        bool fileDone = in.n == 0;
        long long nextSeq = 0, nextToPrint = 0;
        struct Piece {
            char* text = nullptr;
            int64_t textBytes = 0;
            long long cnt[5] = {0, 0, 0, 0, 0};
        };
        std::map<long long, Piece> done;
        std::string err;
        double mapSeconds = 0;
        auto worker = [&](size_t d) {
            void* dImage = nullptr;
            size_t dCap = 0;
            for (;;) {
                dp_record* recs = nullptr;
                int64_t nRecs = 0;
                long long seq;
                {   // cutting is sequential: where a piece ends is only known once it has been split
                    std::unique_lock<std::mutex> lk(mu);
                    if (fileDone || !err.empty()) break;
                    size_t want = pieceBytes;
                    for (;;) {
                        const size_t len = std::min(want, in.n - filePos);
                        const bool final = filePos + len == in.n;
                        if (len > dCap) {
                            dp_device_free(dImage, devices[d]);
                            dImage = nullptr;
                            if (dp_device_alloc(&dImage, len + len / 8, devices[d])) { err = dp_last_error(); break; }
                            dCap = len + len / 8;
                        }
                        int64_t consumed = 0;
                        int fqNext = fq;
                        if (dp_device_copy(dImage, in.p + filePos, len, devices[d])) { err = dp_last_error(); break; }
                        if (dp_split_records((const uint8_t*)dImage, (int64_t)len, minLength, final ? 1 : 0, &fqNext, devices[d], &recs,
                                             &nRecs, &consumed)) {
                            if (!final && strstr(dp_last_error(), "larger piece")) {  // one record longer than the piece
                                want *= 2;
                                continue;
                            }
                            err = dp_last_error();
                            break;
                        }
                        fq = fqNext;
                        filePos += final ? len : (size_t)consumed;
                        if (final) fileDone = true;
                        break;
                    }
                    if (!err.empty()) {
                        cv.notify_all();
                        break;
                    }
                    seq = nextSeq++;
                }
                Piece pc;
                dp_mapping* maps = nullptr;
                int64_t* offs = nullptr;
                const double ta = now_s();
                bool ok = dp_mapper_map_batch_spans(mappers[d], nRecs, (const uint8_t*)dImage, recs, &maps, &offs) == 0;
                const double tb = now_s();
                ok = ok && dp_mapper_paf_block(mappers[d], nRecs, (const uint8_t*)dImage, recs, maps, offs, refName.c_str(), &pc.text,
                                               &pc.textBytes) == 0;
                if (ok)
                    for (int64_t i = 0; i < nRecs; i++) {
                        const int64_t c = offs[i + 1] - offs[i];
                        pc.cnt[4] += recs[i].seq_len;
                        if (c == 1) pc.cnt[0]++;
                        else if (c > 1) pc.cnt[1]++;
                        else pc.cnt[3]++;
                        pc.cnt[2] += c;
                    }
                if (wantStats && ok) {
                    dp_stats st;
                    dp_mapper_get_stats(mappers[d], &st);
                    fprintf(stderr, "[dp_map] gpu %d piece %lld: %lld reads, map %.1f ms, PAF text %.1f ms (%lld bytes)\n", devices[d], seq,
                            (long long)nRecs, (tb - ta) * 1e3, (now_s() - tb) * 1e3, (long long)pc.textBytes);
                }
                dp_free(maps);
                dp_free(offs);
                dp_free(recs);
                {
                    std::lock_guard<std::mutex> lk(mu);
                    if (!ok) err = dp_last_error();
                    mapSeconds += tb - ta;
                    done[seq] = pc;
                }
                cv.notify_all();
                if (!ok) break;
            }
            dp_device_free(dImage, devices[d]);
            cv.notify_all();
        };
        std::vector<std::thread> workers;
        size_t running = devices.size();
        for (size_t d = 0; d < devices.size(); d++)
            workers.emplace_back([&, d] {
                worker(d);
                std::lock_guard<std::mutex> lk(mu);
                running--;
                cv.notify_all();
            });
        long long cnt[5] = {0, 0, 0, 0, 0};
        for (;;) {
            Piece pc;
            {
                std::unique_lock<std::mutex> lk(mu);
                cv.wait(lk, [&] { return done.count(nextToPrint) || running == 0; });
                if (!done.count(nextToPrint)) break;
                pc = done[nextToPrint];
                done.erase(nextToPrint);
                nextToPrint++;
            }
            if (pc.text) fwrite(pc.text, 1, (size_t)pc.textBytes, stdout);
            dp_free(pc.text);
            for (int i = 0; i < 5; i++) cnt[i] += pc.cnt[i];
        }
        for (auto& w : workers) w.join();
        if (!err.empty()) fatal((err.find("Invalid fastq") != std::string::npos ? "" : "dp_map (device io): ") + err);
        fflush(stdout);
        const double t3 = now_s();
        fprintf(stderr, "Uniquely mapped: %lld\nMultiple mappings: %lld\ntotal: %lld\nUnmapped: %lld\n", cnt[0], cnt[1], cnt[2], cnt[3]);
        if (wantStats)
            fprintf(stderr, "[dp_map] %s; device io; gpus=%zu count+values=%.3fs index=%.3fs read+map+print=%.3fs (map calls %.3fs) "
                            "bases=%lld Gbp/s(end to end)=%.3f\n",
                    dp_version(), devices.size(), t1 - t0, t2 - t1, t3 - t2, mapSeconds, cnt[4], cnt[4] / (t3 - t2) / 1e9);
        for (dp_mapper* m : mappers) dp_mapper_destroy(m);
        return 0;
    }

    // ---- reads: a reader thread fills pinned batches; one mapping thread per GPU drains them; output in input order ----
    const size_t nSlots = devices.size() + 2;
    std::vector<Batch> slots(nSlots);
    for (auto& b : slots) {
        b.cap = std::min(batchBytes, in.n + 4096);  // page-locked on first use: a small input pins a small buffer once
        b.reset();
    }
    struct Result {
        dp_mapping* maps = nullptr;
        int64_t* offs = nullptr;
        bool ready = false;
    };
    std::mutex mu;
    std::condition_variable cv;
    std::vector<int> freeSlots, fullSlots;  // slot ids; fullSlots in batch order
    std::vector<long long> slotSeq(nSlots, -1);
    for (size_t i = 0; i < nSlots; i++) freeSlots.push_back((int)i);
    long long produced = 0, nextToMap = 0, nextToPrint = 0;
    bool readerDone = false;
    std::map<long long, std::pair<int, Result>> results;  // batch seq -> (slot, result)
    std::string workerErr;
    double mapSeconds = 0;
    long long totalBases = 0;

    std::thread reader([&] {
        RecordReader rr(in.p, in.n, minLength);
        Record r;
        bool more = rr.next(r);
        while (more) {
            int s;
            {
                std::unique_lock<std::mutex> lk(mu);
                cv.wait(lk, [&] { return !freeSlots.empty() || !workerErr.empty(); });
                if (!workerErr.empty()) break;
                s = freeSlots.back();
                freeSlots.pop_back();
            }
            Batch& b = slots[(size_t)s];
            if (!b.bases && dp_host_alloc((void**)&b.bases, b.cap)) fatal(std::string("dp_host_alloc: ") + dp_last_error());
            b.reset();
            while (more && b.names.size() < batchReads && (b.used + r.len <= b.cap || b.names.empty())) {
                if (r.len > b.cap) fatal("read longer than the batch buffer");
                b.srcs.push_back(r.seq);
                b.used += r.len;
                b.offsets.push_back((int64_t)b.used);
                b.names.emplace_back(r.name, r.nameLen);
                more = rr.next(r);
            }
            // the parser is serial (seqio.go's rules carry state from line to line); the copy into the page-locked
            // batch is not
            parallel_ranges(b.srcs.size(), hostThreads, [&b](size_t, size_t lo, size_t hi) {
                for (size_t i = lo; i < hi; i++)
                    memcpy(b.bases + b.offsets[i], b.srcs[i], (size_t)(b.offsets[i + 1] - b.offsets[i]));
            });
            {
                std::lock_guard<std::mutex> lk(mu);
                slotSeq[(size_t)s] = produced++;
                fullSlots.push_back(s);
            }
            cv.notify_all();
        }
        {
            std::lock_guard<std::mutex> lk(mu);
            readerDone = true;
        }
        cv.notify_all();
    });

    auto mapWorker = [&](size_t d) {
        for (;;) {
            int s;
            long long seq;
            {
                std::unique_lock<std::mutex> lk(mu);
                cv.wait(lk, [&] { return !fullSlots.empty() || readerDone || !workerErr.empty(); });
                if (!workerErr.empty()) return;
                if (fullSlots.empty()) return;
                s = fullSlots.front();
                fullSlots.erase(fullSlots.begin());
                seq = slotSeq[(size_t)s];
                nextToMap++;
            }
            Batch& b = slots[(size_t)s];
            Result res;
            double ta = now_s();
            if (dp_mapper_map_batch(mappers[d], (int64_t)b.names.size(), b.bases, b.offsets.data(), &res.maps, &res.offs)) {
                std::lock_guard<std::mutex> lk(mu);
                workerErr = dp_last_error();
                cv.notify_all();
                return;
            }
            double tb = now_s();
            res.ready = true;
            if (wantStats) {
                dp_stats st;
                dp_mapper_get_stats(mappers[d], &st);
                fprintf(stderr,
                        "[dp_map] gpu %d batch %lld: %zu reads %.1f ms (pack %.1f extract %.1f lookup %.1f chain %.1f host %.1f)\n",
                        devices[d], seq, b.names.size(), (tb - ta) * 1e3, st.ms_pack, st.ms_extract, st.ms_lookup,
                        st.ms_chain, st.ms_host_logic);
            }
            {
                std::lock_guard<std::mutex> lk(mu);
                mapSeconds += tb - ta;
                results[seq] = std::make_pair(s, res);
            }
            cv.notify_all();
        }
    };
    std::vector<std::thread> workers;
    for (size_t d = 0; d < devices.size(); d++) workers.emplace_back(mapWorker, d);

    // ---- printer (commands/map.go:88-106), in batch order ----
    long long mapped = 0, multiple = 0, total = 0, unmapped = 0;
    static char outBuf[1 << 20];
    setvbuf(stdout, outBuf, _IOFBF, sizeof(outBuf));
    for (;;) {
        int s;
        Result res;
        {
            std::unique_lock<std::mutex> lk(mu);
            cv.wait(lk, [&] {
                return results.count(nextToPrint) || !workerErr.empty() ||
                       (readerDone && nextToPrint >= produced);
            });
            if (!workerErr.empty()) break;
            if (!results.count(nextToPrint)) break;  // all batches printed
            s = results[nextToPrint].first;
            res = results[nextToPrint].second;
            results.erase(nextToPrint);
            nextToPrint++;
        }
        Batch& b = slots[(size_t)s];
        // the lines of a batch are formatted by several threads, each into its own buffer, and written in read order
        const size_t nPieces = std::max<size_t>(1, std::min(hostThreads, b.names.size() / 4096 + 1));
        std::vector<std::string> text(nPieces);
        std::vector<std::array<long long, 5>> cnt(nPieces, std::array<long long, 5>{0, 0, 0, 0, 0});
        parallel_ranges(b.names.size(), nPieces, [&](size_t t, size_t lo, size_t hi) {
            std::vector<char> line(1 << 16);
            std::string nameBuf;
            std::string& out = text[t];
            out.reserve((hi - lo) * 96);
            for (size_t i = lo; i < hi; i++) {
                const int64_t a = res.offs[i], e = res.offs[i + 1];
                const int64_t qlen = b.offsets[i + 1] - b.offsets[i];
                cnt[t][4] += qlen;
                if (e > a) {
                    nameBuf.assign((const char*)b.names[i].first, b.names[i].second);
                    for (int64_t j = a; j < e; j++) {
                        int n = dp_mapper_paf_line(mappers[0], res.maps + j, nameBuf.c_str(), qlen, refName.c_str(),
                                                   line.data(), (int)line.size());
                        if (n < 0) fatal("PAF line too long");
                        out.append(line.data(), (size_t)n);
                        out.push_back('\n');
                    }
                    if (e - a == 1) cnt[t][0]++;
                    else cnt[t][1]++;
                    cnt[t][2] += e - a;
                } else {
                    cnt[t][3]++;
                }
            }
        });
        for (size_t t = 0; t < nPieces; t++) {
            fwrite(text[t].data(), 1, text[t].size(), stdout);
            mapped += cnt[t][0];
            multiple += cnt[t][1];
            total += cnt[t][2];
            unmapped += cnt[t][3];
            totalBases += cnt[t][4];
        }
        dp_free(res.maps);
        dp_free(res.offs);
        {
            std::lock_guard<std::mutex> lk(mu);
            freeSlots.push_back(s);
        }
        cv.notify_all();
    }
    reader.join();
    for (auto& w : workers) w.join();
    if (!workerErr.empty()) fatal("dp_mapper_map_batch: " + workerErr);
    fflush(stdout);
    double t3 = now_s();
    fprintf(stderr, "Uniquely mapped: %lld\nMultiple mappings: %lld\ntotal: %lld\nUnmapped: %lld\n", mapped, multiple, total,
            unmapped);
    if (wantStats) {
        int64_t info[5] = {0, 0, 0, 0, 0};
        dp_mapper_index_info(mappers[0], info);
        fprintf(stderr,
                "[dp_map] %s; gpus=%zu seeds=%lld chunks=%lld count+values=%.3fs index=%.3fs read+map+print=%.3fs "
                "(map calls %.3fs) bases=%lld Gbp/s(end to end)=%.3f\n",
                dp_version(), devices.size(), (long long)info[0], (long long)info[1], t1 - t0, t2 - t1, t3 - t2, mapSeconds,
                totalBases, totalBases / (t3 - t2) / 1e9);
    }
    for (auto& b : slots)
        if (b.bases) dp_host_free(b.bases);
    for (dp_mapper* m : mappers) dp_mapper_destroy(m);
    return 0;
}
