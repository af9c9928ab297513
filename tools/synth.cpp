// Seeded synthetic data for tests and bench.py: i.i.d. uniform reference genomes and ONT-like reads
// (fixed length, uniform start, fair strand coin, per-base substitution / insertion / deletion).
// PRNG: xoshiro256** seeded through splitmix64. Every read has its own stream derived from (seed, read index), so
// any shard of the read set can be generated independently (ranks, threads) with identical bytes.
#include <cstdint>
#include <cstring>
#include <thread>
#include <vector>

namespace {
struct SplitMix64 {
    uint64_t s;
    explicit SplitMix64(uint64_t seed) : s(seed) {}
    uint64_t next() {
        uint64_t z = (s += 0x9E3779B97F4A7C15ULL);
        z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
        z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
        return z ^ (z >> 31);
    }
};
struct Xoshiro256ss {
    uint64_t s[4];
    explicit Xoshiro256ss(uint64_t seed) {
        SplitMix64 sm(seed);
        for (int i = 0; i < 4; i++) s[i] = sm.next();
    }
    static uint64_t rotl(uint64_t x, int k) { return (x << k) | (x >> (64 - k)); }
    uint64_t next() {
        uint64_t result = rotl(s[1] * 5, 7) * 9;
        uint64_t t = s[1] << 17;
        s[2] ^= s[0];
        s[3] ^= s[1];
        s[1] ^= s[2];
        s[0] ^= s[3];
        s[2] ^= t;
        s[3] = rotl(s[3], 45);
        return result;
    }
    double uniform() { return (double)(next() >> 11) * (1.0 / 9007199254740992.0); }
    uint64_t below(uint64_t n) { return (uint64_t)(((unsigned __int128)next() * n) >> 64); }
};
const char kBases[4] = {'A', 'C', 'G', 'T'};
inline int code_of(char c) { return ((c >> 1) ^ ((c & 4) >> 2)) & 3; }

void gen_reference_range(uint64_t seed, int64_t begin, int64_t end, char* out) {
    // 32 bases per PRNG draw; block b covers bases [32b, 32b+32) and has its own position in the stream
    const int64_t BLK = 1 << 16;  // bases per independently seeded block
    for (int64_t b0 = begin; b0 < end;) {
        int64_t blk = b0 / BLK;
        Xoshiro256ss rng(seed * 0x9E3779B97F4A7C15ULL + (uint64_t)blk + 1);
        int64_t bstart = blk * BLK;
        int64_t bend = bstart + BLK;
        for (int64_t p = bstart; p < bend && p < end; p += 32) {
            uint64_t r = rng.next();
            for (int j = 0; j < 32; j++) {
                int64_t q = p + j;
                if (q >= b0 && q < end) out[q] = kBases[(r >> (2 * j)) & 3];
            }
        }
        b0 = bend;
    }
}
}  // namespace

extern "C" {

// Uniform i.i.d. reference of `len` bases into out[0..len).
void dps_reference(uint64_t seed, int64_t len, char* out, int threads) {
    if (threads < 1) threads = 1;
    const int64_t BLK = 1 << 16;
    int64_t nblk = (len + BLK - 1) / BLK;
    std::vector<std::thread> th;
    for (int t = 0; t < threads; t++) {
        int64_t b0 = nblk * t / threads * BLK, b1 = nblk * (t + 1) / threads * BLK;
        if (b1 > len) b1 = len;
        if (b0 >= b1) continue;
        th.emplace_back(gen_reference_range, seed, b0, b1, out);
    }
    for (auto& t : th) t.join();
}

// The `rep` variant of a reference: about `frac` of the bases of ref[0..len) are overwritten in place by
// copies of `families` random repeat units of min_len..max_len bases, every copy with its own substitution divergence
// drawn uniformly from [0, max_div] and a fair strand coin. Serial and seeded: identical bytes everywhere.
void dps_make_repeats(char* ref, int64_t len, uint64_t seed, int families, double frac, int64_t min_len, int64_t max_len,
                      double max_div) {
    Xoshiro256ss rng(seed * 0xA24BAED4963EE407ULL + 99);
    std::vector<std::vector<char>> fam((size_t)families);
    for (auto& f : fam) {
        int64_t L = min_len + (int64_t)rng.below((uint64_t)(max_len - min_len + 1));
        f.resize((size_t)L);
        for (auto& c : f) c = kBases[rng.below(4)];
    }
    const int64_t target = (int64_t)(frac * (double)len);
    int64_t placed = 0;
    while (placed < target) {
        const std::vector<char>& f = fam[(size_t)rng.below((uint64_t)families)];
        const int64_t L = (int64_t)f.size();
        if (L >= len) break;
        const int64_t pos = (int64_t)rng.below((uint64_t)(len - L));
        const double div = rng.uniform() * max_div;
        const int strand = (int)(rng.next() >> 63);
        for (int64_t j = 0; j < L; j++) {
            char b = strand ? kBases[3 - code_of(f[(size_t)(L - 1 - j)])] : f[(size_t)j];
            if (rng.uniform() < div) b = kBases[(code_of(b) + 1 + (int)rng.below(3)) & 3];
            ref[pos + j] = b;
        }
        placed += L;
    }
}

// Reads first_index .. first_index+n-1 of the read set `seed`, each exactly read_len bases, concatenated into `out`.
// truth (optional, 2 int64 per read): template start on the reference, strand (0 '+', 1 '-').
void dps_reads(const char* ref, int64_t ref_len, int circular, uint64_t seed, int64_t first_index, int64_t n,
               int64_t read_len, double p_sub, double p_ins, double p_del, char* out, int64_t* truth, int threads) {
    if (threads < 1) threads = 1;
    auto work = [&](int64_t i0, int64_t i1) {
        std::vector<char> tmp((size_t)read_len);
        for (int64_t i = i0; i < i1; i++) {
            uint64_t idx = (uint64_t)(first_index + i);
            Xoshiro256ss rng(seed * 0xD1342543DE82EF95ULL + idx * 0x9E3779B97F4A7C15ULL + 12345);
            int64_t span = read_len + read_len / 4 + 64;  // template bases a read may consume
            int64_t start;
            if (circular || ref_len <= span) start = (int64_t)rng.below((uint64_t)ref_len);
            else start = (int64_t)rng.below((uint64_t)(ref_len - span));
            int strand = (int)(rng.next() >> 63);
            int64_t p = start;
            int64_t m = 0;
            while (m < read_len) {
                double r = rng.uniform();
                char tb;
                if (p >= ref_len) {
                    if (circular) {
                        p -= ref_len;
                        tb = ref[p];
                    } else {
                        tb = kBases[rng.below(4)];
                    }
                } else {
                    tb = ref[p];
                }
                if (r < p_ins) {
                    tmp[(size_t)m++] = kBases[rng.below(4)];
                } else if (r < p_ins + p_del) {
                    p++;
                } else if (r < p_ins + p_del + p_sub) {
                    tmp[(size_t)m++] = kBases[(code_of(tb) + 1 + (int)rng.below(3)) & 3];
                    p++;
                } else {
                    tmp[(size_t)m++] = tb;
                    p++;
                }
            }
            char* o = out + i * read_len;
            if (!strand) {
                memcpy(o, tmp.data(), (size_t)read_len);
            } else {
                for (int64_t j = 0; j < read_len; j++) o[j] = kBases[3 - code_of(tmp[(size_t)(read_len - 1 - j)])];
            }
            if (truth) {
                truth[2 * i] = start;
                truth[2 * i + 1] = strand;
            }
        }
    };
    std::vector<std::thread> th;
    for (int t = 0; t < threads; t++) {
        int64_t i0 = n * t / threads, i1 = n * (t + 1) / threads;
        if (i0 >= i1) continue;
        th.emplace_back(work, i0, i1);
    }
    for (auto& t : th) t.join();
}

}  // extern "C"
