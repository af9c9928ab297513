// ORACLE — TEST INFRASTRUCTURE ONLY (see oracle.hpp).
// `oracle_map`: CPU twin of `downpore map` (commands/map.go:33-116) over the oracle. Output: PAF lines grouped per
// read in input order (canonical order for the reference's scheduling-dependent line order), stats on stderr.
#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <string>
#include <thread>
#include <vector>

#include "oracle.hpp"

using namespace dpo;

int main(int argc, char** argv) {
    // downpore.go:34-51 parseArgs: `-name value` pairs, any number of leading dashes
    std::map<std::string, std::string> args = {{"input", ""},        {"reference", ""},   {"circular", "true"},
                                               {"k", "11"},          {"query_size", "1000"}, {"min_length", "500"},
                                               {"chunk_size", "10000"}, {"seed_rate", "40"}, {"num_workers", "4"}};
    std::map<std::string, std::string> alias = {{"i", "input"},      {"r", "reference"}, {"ci", "circular"},
                                                {"k", "k"},          {"q", "query_size"}, {"m", "min_length"},
                                                {"ch", "chunk_size"}, {"s", "seed_rate"}, {"n", "num_workers"}};
    for (int i = 1; i + 1 < argc; i += 2) {
        std::string name = argv[i];
        while (!name.empty() && name[0] == '-') name.erase(0, 1);
        if (alias.count(name)) name = alias[name];
        if (!args.count(name)) {
            fprintf(stderr, "Unknown argument: %s\n", argv[i]);
            return 2;
        }
        args[name] = argv[i + 1];
    }
    try {
        gint k = atoll(args["k"].c_str());
        int numWorkers = atoi(args["num_workers"].c_str());
        gint minLength = atoll(args["min_length"].c_str());
        const std::string& cs = args["circular"];
        bool circular = cs == "1" || (!cs.empty() && (cs[0] == 'T' || cs[0] == 't'));  // command.go:72-74
        gint querySize = atoll(args["query_size"].c_str());
        gint chunkSize = atoll(args["chunk_size"].c_str());
        gint seedRate = atoll(args["seed_rate"].c_str());

        auto t0 = std::chrono::steady_clock::now();
        std::vector<FastaRecord> refs = ReadFasta(args["reference"], 0);
        if (refs.empty()) {
            fprintf(stderr, "no reference sequence\n");
            return 1;
        }
        auto refName = std::make_shared<std::string>(refs[0].name);
        PackedSeq reference = NewPackedSequence(0, refs[0].seq, refName);
        std::vector<uint64_t> counts;
        for (size_t i = 0; i < refs.size(); i++) {  // Q14: counts cover all records
            PackedSeq r = NewPackedSequence((gint)i, refs[i].seq, nullptr);
            KmerOccurrences(r, k, counts);
        }
        std::vector<double> values = KmerValues(counts, k);
        fprintf(stderr, "K-mer counting complete. Preparing to start indexing and querying...\n");
        auto t1 = std::chrono::steady_clock::now();
        Mapper mapper;
        NewMapper(mapper, reference, circular, k, values.data(), seedRate, querySize, chunkSize);
        mapper.refName = *refName;
        auto t2 = std::chrono::steady_clock::now();
        std::vector<FastaRecord> reads = ReadFasta(args["input"], minLength);
        auto t3 = std::chrono::steady_clock::now();
        std::vector<std::vector<Mapping>> res(reads.size());
        std::atomic<size_t> next(0);
        std::vector<Counters> ctr((size_t)std::max(1, numWorkers));
        auto work = [&](int t) {
            for (;;) {
                size_t i = next.fetch_add(1);
                if (i >= reads.size()) break;
                PackedSeq q = NewPackedSequence((gint)i, reads[i].seq, nullptr);
                res[i] = Map(mapper, q, &ctr[(size_t)t]);
            }
        };
        std::vector<std::thread> th;
        for (int t = 1; t < numWorkers; t++) th.emplace_back(work, t);
        work(0);
        for (auto& t : th) t.join();
        auto t4 = std::chrono::steady_clock::now();
        long long mapped = 0, multiple = 0, total = 0, unmapped = 0, bases = 0;
        for (size_t i = 0; i < reads.size(); i++) {
            bases += (long long)reads[i].seq.size();
            if (!res[i].empty()) {
                for (const Mapping& m : res[i]) puts(AsString(mapper, m, reads[i].name).c_str());
                if (res[i].size() == 1) mapped++;
                else multiple++;
                total += (long long)res[i].size();
            } else {
                unmapped++;
            }
        }
        fprintf(stderr, "Uniquely mapped: %lld\nMultiple mappings: %lld\ntotal: %lld\nUnmapped: %lld\n", mapped, multiple,
                total, unmapped);
        auto sec = [](auto a, auto b) { return std::chrono::duration<double>(b - a).count(); };
        Counters tot;
        for (auto& c : ctr) tot.add(c);
        fprintf(stderr,
                "[oracle] seeds=%lld chunks=%zu count+values=%.3fs index=%.3fs read=%.3fs map=%.3fs (%d threads) "
                "bases=%lld Gbp/s=%.4f\n",
                mapper.index.size, mapper.index.sequences.size(), sec(t0, t1), sec(t1, t2), sec(t2, t3), sec(t3, t4),
                numWorkers, bases, bases / sec(t3, t4) / 1e9);
        fprintf(stderr,
                "[oracle] windows=%lld kmer_lookups=%lld query_seeds=%lld posting_runs=%lld posting_entries=%lld "
                "candidates=%lld cand_pass=%lld chain_cells=%lld chains=%lld sort_ties_unpinned=%lld\n",
                tot.windows, tot.kmer_lookups, tot.query_seeds, tot.posting_runs, tot.posting_entries, tot.candidates,
                tot.cand_pass, tot.chain_cells, tot.chains, tot.sort_ties_unpinned);
    } catch (const std::exception& ex) {
        fprintf(stderr, "fatal: %s\n", ex.what());
        return 1;
    }
    return 0;
}
