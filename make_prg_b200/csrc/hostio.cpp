// Host-side I/O either side of the hot path, plain C++ on host threads:
//
//   loader   load_alignment_file                 make_prg/utils/io_utils.py:17-49
//            (FASTA / FASTA.gz -> upper-cased row-major ASCII matrices, ready for mprg_build_ascii)
//   writers  PrgEncoder.encode / write           make_prg/utils/prg_encoder.py:44-91
//            GFA_Output.build_gfa_string         make_prg/utils/gfa.py:39-109
//            InputOutputFiles.create_final_files make_prg/utils/input_output_files.py:70-135
//            (.prg.fa sorted by file name, .prg.bin / .prg.gfa for one locus, stored .zip archives
//            of <locus>.bin / <locus>.gfa for several -- zipfile.ZipFile's default ZIP_STORED)
//
// Nothing here touches the GPU except the optional pinned allocation of the loader's output buffer.
#include <cuda_runtime.h>
#include <errno.h>
#include <fcntl.h>
#include <string.h>
#include <stdlib.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <time.h>
#include <unistd.h>
#include <zlib.h>

#include <immintrin.h>

#include <algorithm>
#include <memory>
#include <mutex>
#include <sched.h>
#include <atomic>
#include <string>
#include <thread>
#include <unordered_map>
#include <vector>

#include "../../include/mprg.h"

namespace {

template <typename F>
void parallel_for_t(int n, int n_threads, F fn) {  // fn(item, thread index)
    n_threads = std::max(1, std::min(n_threads, n));
    if (n_threads == 1) {
        for (int i = 0; i < n; ++i) fn(i, 0);
        return;
    }
    std::atomic<int> next{0};
    std::vector<std::thread> th;
    for (int t = 0; t < n_threads; ++t)
        th.emplace_back([&, t]() {
            for (int i = next.fetch_add(1); i < n; i = next.fetch_add(1)) fn(i, t);
        });
    for (auto &t : th) t.join();
}

template <typename F>
void parallel_for(int n, int n_threads, F fn) {
    n_threads = std::max(1, std::min(n_threads, n));
    if (n_threads == 1) {
        for (int i = 0; i < n; ++i) fn(i);
        return;
    }
    std::atomic<int> next{0};
    std::vector<std::thread> th;
    for (int t = 0; t < n_threads; ++t)
        th.emplace_back([&]() {
            for (int i = next.fetch_add(1); i < n; i = next.fetch_add(1)) fn(i);
        });
    for (auto &t : th) t.join();
}

// ------------------------------------------------------------------------------------------------
// loader
// ------------------------------------------------------------------------------------------------
// Big host buffers come in 2 MB-aligned blocks advised as transparent huge pages: first-touch page
// faults are what a loader of fresh memory otherwise spends most of its time in.
uint8_t *huge_alloc(size_t bytes) {
    const size_t two_mb = (size_t)2 << 20;
    bytes = (bytes + two_mb - 1) & ~(two_mb - 1);
    void *p = aligned_alloc(two_mb, bytes);
    if (p) madvise(p, bytes, MADV_HUGEPAGE);
    return (uint8_t *)p;
}

// Pinned host buffers are expensive to create and to destroy (tens of ms per 100 MB): the loader's
// output buffers are recycled through a small process-wide pool (at most MPRG_PINNED_POOL_MB, default
// 1024, are kept when idle).
struct PinnedPool {
    std::mutex m;
    std::vector<std::pair<uint8_t *, size_t>> idle;
    size_t idle_bytes = 0;
    uint8_t *acquire(size_t bytes, size_t &capacity) {
        {
            std::lock_guard<std::mutex> lock(m);
            int best = -1;
            for (int i = 0; i < (int)idle.size(); ++i)
                if (idle[i].second >= bytes && (best < 0 || idle[i].second < idle[best].second)) best = i;
            if (best >= 0 && idle[best].second <= 4 * bytes + ((size_t)16 << 20)) {
                uint8_t *p = idle[best].first;
                capacity = idle[best].second;
                idle_bytes -= capacity;
                idle.erase(idle.begin() + best);
                return p;
            }
        }
        capacity = ((bytes + bytes / 8) + ((size_t)8 << 20) - 1) & ~(((size_t)8 << 20) - 1);
        uint8_t *p = nullptr;
        if (cudaHostAlloc((void **)&p, capacity, cudaHostAllocDefault) != cudaSuccess) {
            cudaGetLastError();
            return nullptr;
        }
        return p;
    }
    void release(uint8_t *p, size_t capacity) {
        static const size_t limit = []() {
            const char *e = getenv("MPRG_PINNED_POOL_MB");
            return (size_t)(e ? atoll(e) : 1024) << 20;
        }();
        {
            std::lock_guard<std::mutex> lock(m);
            if (idle_bytes + capacity <= limit) {
                idle.emplace_back(p, capacity);
                idle_bytes += capacity;
                return;
            }
        }
        cudaFreeHost(p);
    }
};
PinnedPool &pinned_pool() {
    static PinnedPool *pool = new PinnedPool();  // never destroyed: no CUDA calls at process exit
    return *pool;
}

// 32 MB blocks of ordinary host memory (huge-page advised) for the loader's parse and read buffers, kept
// between calls: fresh pages cost a fault and a zero-fill each, a loader call touches twice its input
// in scratch, and its threads do not outlive it.  At most MPRG_HOST_POOL_MB (default 1024) stay idle.
struct BlockPool {
    static constexpr size_t BLOCK = (size_t)32 << 20;
    std::mutex m;
    std::vector<uint8_t *> idle;
    uint8_t *take() {
        {
            std::lock_guard<std::mutex> lock(m);
            if (!idle.empty()) {
                uint8_t *p = idle.back();
                idle.pop_back();
                return p;
            }
        }
        return huge_alloc(BLOCK);
    }
    void give(uint8_t *p) {
        static const size_t limit = []() {
            const char *e = getenv("MPRG_HOST_POOL_MB");
            return (size_t)(e ? atoll(e) : 1024) << 20;
        }();
        {
            std::lock_guard<std::mutex> lock(m);
            if ((idle.size() + 1) * BLOCK <= limit) {
                idle.push_back(p);
                return;
            }
        }
        free(p);
    }
};
BlockPool &block_pool() {
    static BlockPool *pool = new BlockPool();
    return *pool;
}

// bump allocator of one parsing thread; the blocks live until the matrices have been copied out
struct Slab {
    static constexpr size_t BLOCK = BlockPool::BLOCK;
    std::vector<uint8_t *> blocks;  // pool blocks, newest last
    std::vector<uint8_t *> big;     // loci that need more than a quarter of a block: own allocation
    size_t used = BLOCK;
    uint8_t *take(size_t bytes) {
        if (bytes > BLOCK / 4) {
            uint8_t *p = huge_alloc(bytes);
            if (p) big.push_back(p);
            return p;
        }
        if (used + bytes > BLOCK) {
            uint8_t *p = block_pool().take();
            if (!p) return nullptr;
            blocks.push_back(p);
            used = 0;
        }
        uint8_t *p = blocks.back() + used;
        used += (bytes + 63) & ~(size_t)63;
        return p;
    }
    void give_back(uint8_t *p, size_t taken, size_t kept) {  // shrink the newest allocation
        if (!blocks.empty() && p >= blocks.back() && p + ((taken + 63) & ~(size_t)63) == blocks.back() + used)
            used = (size_t)(p - blocks.back()) + ((kept + 63) & ~(size_t)63);
    }
    ~Slab() {
        for (uint8_t *b : blocks) block_pool().give(b);
        for (uint8_t *b : big) free(b);
    }
};

struct ParsedFile {
    uint8_t *buf = nullptr;  // the matrix, row-major (in the parsing thread's slab)
    std::string titles;        // header lines without '>', joined by '\n'
    int64_t matrix_bytes = 0;
    int32_t n_rows = 0, n_cols = 0, status = MPRG_LOAD_OK, flags = 0;
};

bool ends_with(const char *s, const char *suffix) {
    const size_t n = strlen(s), m = strlen(suffix);
    return n >= m && memcmp(s + n - m, suffix, m) == 0;
}

// Reads the file into `block` (capacity bytes) when it fits, else into `big`; *text / *size = where it is.
bool read_plain(const char *path, uint8_t *block, size_t capacity, std::vector<uint8_t> &big, const uint8_t **text,
                size_t *size) {
    const int fd = open(path, O_RDONLY);
    if (fd < 0) return false;
    struct stat st;
    if (fstat(fd, &st) != 0 || !S_ISREG(st.st_mode)) {
        close(fd);
        return false;
    }
    const size_t want = (size_t)st.st_size;
    uint8_t *dst = block;
    if (!block || want > capacity) {
        big.resize(want);
        dst = big.data();
    }
    size_t got = 0;
    while (got < want) {
        const ssize_t r = read(fd, dst + got, want - got);
        if (r < 0) {
            if (errno == EINTR) continue;
            close(fd);
            return false;
        }
        if (r == 0) break;
        got += (size_t)r;
    }
    close(fd);
    *text = dst;
    *size = got;
    return true;
}

bool read_gzip(const char *path, std::vector<uint8_t> &out) {
    gzFile f = gzopen(path, "rb");
    if (!f) return false;
    gzbuffer(f, 1 << 20);
    out.clear();
    size_t cap = 1 << 22;
    out.resize(cap);
    size_t got = 0;
    while (true) {
        if (got == cap) {
            cap *= 2;
            out.resize(cap);
        }
        const int r = gzread(f, out.data() + got, (unsigned)std::min<size_t>(cap - got, 1u << 30));
        if (r < 0) {
            gzclose(f);
            return false;
        }
        if (r == 0) break;
        got += (size_t)r;
    }
    gzclose(f);
    out.resize(got);
    return true;
}

// One text line as Python's universal-newline reader cuts it: '\n', '\r\n' and a lone '\r' all end it.
inline void next_line(const uint8_t *s, const uint8_t *end, const uint8_t *&line_end, const uint8_t *&next) {
    const uint8_t *p = (const uint8_t *)memchr(s, '\n', (size_t)(end - s));
    const uint8_t *lim = p ? p : end;
    const uint8_t *q = (const uint8_t *)memchr(s, '\r', (size_t)(lim - s));
    if (q) {
        line_end = q;
        next = q + 1;
        if (next < end && *next == '\n') ++next;
    } else {
        line_end = lim;
        next = p ? p + 1 : end;
    }
}

// The sequence lines of one record: copies [s, ...) to w up to (not including) the next title line
// (a '>' first in its line) or `end`, dropping line ends and blanks and upper-casing a-z.
// bits |= 1 when a byte >= 0x80 was copied, |= 2 when an 'N' was.  Returns the read position.
// Strict UTF-8 as Python's decoder accepts it (no overlong forms, no surrogates, nothing above U+10FFFF).
bool valid_utf8(const uint8_t *p, size_t n) {
    size_t i = 0;
    while (i < n) {
        const uint8_t c = p[i];
        if (c < 0x80) {
            ++i;
            continue;
        }
        int len;
        uint32_t cp, lo;
        if ((c & 0xE0) == 0xC0) { len = 2; cp = c & 0x1F; lo = 0x80; }
        else if ((c & 0xF0) == 0xE0) { len = 3; cp = c & 0x0F; lo = 0x800; }
        else if ((c & 0xF8) == 0xF0) { len = 4; cp = c & 0x07; lo = 0x10000; }
        else return false;
        if (i + (size_t)len > n) return false;
        for (int k = 1; k < len; ++k) {
            if ((p[i + k] & 0xC0) != 0x80) return false;
            cp = (cp << 6) | (p[i + k] & 0x3F);
        }
        if (cp < lo || cp > 0x10FFFF || (cp >= 0xD800 && cp <= 0xDFFF)) return false;
        i += (size_t)len;
    }
    return true;
}

// What str.rstrip() removes at the end of a line besides blanks and line ends (Biopython's
// SimpleFastaParser rstrips every line before joining them): \t \v \f and the separators 0x1c-0x1f.
inline bool is_trailing_ws(uint8_t c) { return c == '\t' || c == 0x0b || c == 0x0c || (c >= 0x1c && c <= 0x1f); }
inline void rstrip_line(uint8_t *row_start, uint8_t *&w) {
    while (w > row_start && is_trailing_ws(w[-1])) --w;
}

const uint8_t *copy_record_scalar(const uint8_t *s, const uint8_t *end, uint8_t *&w, int &bits, uint8_t *row_start) {
    unsigned hi = 0, n_count = 0;
    while (s < end) {
        uint8_t c = *s++;
        if (c == '\n' || c == '\r') {
            rstrip_line(row_start, w);
            if (c == '\r' && s < end && *s == '\n') ++s;
            if (s < end && *s == '>') break;
            continue;
        }
        if (c == ' ') continue;
        c = (uint8_t)(c - (((uint8_t)(c - 'a') < 26) << 5));
        hi |= c;
        n_count += (c == 'N');
        *w++ = c;
    }
    rstrip_line(row_start, w);  // a last line without a line end
    bits |= ((hi & 0x80u) ? 1 : 0) | (n_count ? 2 : 0);
    return s;
}

alignas(32) const uint8_t LANE_MASK[64] = {
    255, 255, 255, 255, 255, 255, 255, 255, 255, 255, 255, 255, 255, 255, 255, 255,
    255, 255, 255, 255, 255, 255, 255, 255, 255, 255, 255, 255, 255, 255, 255, 255};

// The same, 32 bytes per step (the output buffer has 32 bytes of slack: whole vectors are stored and
// the write position advances by the bytes that count).
__attribute__((target("avx2"))) const uint8_t *copy_record_avx2(const uint8_t *s, const uint8_t *end, uint8_t *&w,
                                                                  int &bits, uint8_t *row_start) {
    const __m256i v_nl = _mm256_set1_epi8('\n'), v_cr = _mm256_set1_epi8('\r'), v_sp = _mm256_set1_epi8(' ');
    const __m256i v_lo = _mm256_set1_epi8('a' - 1), v_hi = _mm256_set1_epi8('z' + 1);
    const __m256i v_case = _mm256_set1_epi8(0x20), v_n = _mm256_set1_epi8('N');
    __m256i acc_hi = _mm256_setzero_si256(), acc_n = _mm256_setzero_si256();
    while (s + 32 <= end) {
        const __m256i v = _mm256_loadu_si256((const __m256i *)s);
        const __m256i special = _mm256_or_si256(_mm256_or_si256(_mm256_cmpeq_epi8(v, v_nl), _mm256_cmpeq_epi8(v, v_cr)),
                                                _mm256_cmpeq_epi8(v, v_sp));
        const __m256i lower = _mm256_and_si256(_mm256_cmpgt_epi8(v, v_lo), _mm256_cmpgt_epi8(v_hi, v));
        const __m256i up = _mm256_xor_si256(v, _mm256_and_si256(lower, v_case));
        const uint32_t m = (uint32_t)_mm256_movemask_epi8(special);
        _mm256_storeu_si256((__m256i *)w, up);
        if (m == 0) {
            acc_hi = _mm256_or_si256(acc_hi, v);
            acc_n = _mm256_or_si256(acc_n, _mm256_cmpeq_epi8(up, v_n));
            w += 32;
            s += 32;
            continue;
        }
        const int k = __builtin_ctz(m);
        const __m256i lanes = _mm256_loadu_si256((const __m256i *)(LANE_MASK + 32 - k));
        acc_hi = _mm256_or_si256(acc_hi, _mm256_and_si256(v, lanes));
        acc_n = _mm256_or_si256(acc_n, _mm256_and_si256(_mm256_cmpeq_epi8(up, v_n), lanes));
        w += k;
        const uint8_t c = s[k];
        s += k + 1;
        if (c == ' ') continue;
        rstrip_line(row_start, w);
        if (c == '\r' && s < end && *s == '\n') ++s;
        if (s < end && *s == '>') {
            bits |= (_mm256_movemask_epi8(acc_hi) ? 1 : 0) | (_mm256_movemask_epi8(acc_n) ? 2 : 0);
            return s;
        }
    }
    bits |= (_mm256_movemask_epi8(acc_hi) ? 1 : 0) | (_mm256_movemask_epi8(acc_n) ? 2 : 0);
    return copy_record_scalar(s, end, w, bits, row_start);
}

// ------------------------------------------------------------------------------------------------
// N replacement (io_utils.py:35-47, seq_utils.py:246-290): every N becomes the majority symbol of its
// column among the non-gap, non-N symbols; ties and empty columns are resolved by a private
// random.Random seeded with sha256 of the concatenated rows, one `choice` per column in column order.
// CPython's random.seed(bytes) (version 2) and random.choice are restated here: SHA-256 / SHA-512,
// MT19937 init_by_array, getrandbits-based _randbelow.
// ------------------------------------------------------------------------------------------------
struct Sha256 {
    uint32_t h[8] = {0x6a09e667u, 0xbb67ae85u, 0x3c6ef372u, 0xa54ff53au, 0x510e527fu, 0x9b05688cu, 0x1f83d9abu, 0x5be0cd19u};
    static inline uint32_t rotr(uint32_t x, int n) { return (x >> n) | (x << (32 - n)); }
    void block(const uint8_t *p) {
        static const uint32_t K[64] = {
            0x428a2f98u, 0x71374491u, 0xb5c0fbcfu, 0xe9b5dba5u, 0x3956c25bu, 0x59f111f1u, 0x923f82a4u, 0xab1c5ed5u,
            0xd807aa98u, 0x12835b01u, 0x243185beu, 0x550c7dc3u, 0x72be5d74u, 0x80deb1feu, 0x9bdc06a7u, 0xc19bf174u,
            0xe49b69c1u, 0xefbe4786u, 0x0fc19dc6u, 0x240ca1ccu, 0x2de92c6fu, 0x4a7484aau, 0x5cb0a9dcu, 0x76f988dau,
            0x983e5152u, 0xa831c66du, 0xb00327c8u, 0xbf597fc7u, 0xc6e00bf3u, 0xd5a79147u, 0x06ca6351u, 0x14292967u,
            0x27b70a85u, 0x2e1b2138u, 0x4d2c6dfcu, 0x53380d13u, 0x650a7354u, 0x766a0abbu, 0x81c2c92eu, 0x92722c85u,
            0xa2bfe8a1u, 0xa81a664bu, 0xc24b8b70u, 0xc76c51a3u, 0xd192e819u, 0xd6990624u, 0xf40e3585u, 0x106aa070u,
            0x19a4c116u, 0x1e376c08u, 0x2748774cu, 0x34b0bcb5u, 0x391c0cb3u, 0x4ed8aa4au, 0x5b9cca4fu, 0x682e6ff3u,
            0x748f82eeu, 0x78a5636fu, 0x84c87814u, 0x8cc70208u, 0x90befffau, 0xa4506cebu, 0xbef9a3f7u, 0xc67178f2u};
        uint32_t w[64];
        for (int i = 0; i < 16; ++i)
            w[i] = ((uint32_t)p[4 * i] << 24) | ((uint32_t)p[4 * i + 1] << 16) | ((uint32_t)p[4 * i + 2] << 8) | p[4 * i + 3];
        for (int i = 16; i < 64; ++i) {
            const uint32_t s0 = rotr(w[i - 15], 7) ^ rotr(w[i - 15], 18) ^ (w[i - 15] >> 3);
            const uint32_t s1 = rotr(w[i - 2], 17) ^ rotr(w[i - 2], 19) ^ (w[i - 2] >> 10);
            w[i] = w[i - 16] + s0 + w[i - 7] + s1;
        }
        uint32_t a = h[0], b = h[1], c = h[2], d = h[3], e = h[4], f = h[5], g = h[6], hh = h[7];
        for (int i = 0; i < 64; ++i) {
            const uint32_t t1 = hh + (rotr(e, 6) ^ rotr(e, 11) ^ rotr(e, 25)) + ((e & f) ^ (~e & g)) + K[i] + w[i];
            const uint32_t t2 = (rotr(a, 2) ^ rotr(a, 13) ^ rotr(a, 22)) + ((a & b) ^ (a & c) ^ (b & c));
            hh = g; g = f; f = e; e = d + t1; d = c; c = b; b = a; a = t1 + t2;
        }
        h[0] += a; h[1] += b; h[2] += c; h[3] += d; h[4] += e; h[5] += f; h[6] += g; h[7] += hh;
    }
    void digest(const uint8_t *data, uint64_t n, uint8_t out[32]) {
        uint64_t i = 0;
        for (; i + 64 <= n; i += 64) block(data + i);
        uint8_t tail[128] = {0};
        const uint64_t rem = n - i;
        memcpy(tail, data + i, (size_t)rem);
        tail[rem] = 0x80;
        const size_t total = rem + 9 <= 64 ? 64 : 128;
        const uint64_t bits = n * 8;
        for (int k = 0; k < 8; ++k) tail[total - 1 - k] = (uint8_t)(bits >> (8 * k));
        block(tail);
        if (total == 128) block(tail + 64);
        for (int k = 0; k < 8; ++k)
            for (int j = 0; j < 4; ++j) out[4 * k + j] = (uint8_t)(h[k] >> (24 - 8 * j));
    }
};

struct Sha512 {
    uint64_t h[8] = {0x6a09e667f3bcc908ull, 0xbb67ae8584caa73bull, 0x3c6ef372fe94f82bull, 0xa54ff53a5f1d36f1ull,
                     0x510e527fade682d1ull, 0x9b05688c2b3e6c1full, 0x1f83d9abfb41bd6bull, 0x5be0cd19137e2179ull};
    static inline uint64_t rotr(uint64_t x, int n) { return (x >> n) | (x << (64 - n)); }
    void block(const uint8_t *p) {
        static const uint64_t K[80] = {
            0x428a2f98d728ae22ull, 0x7137449123ef65cdull, 0xb5c0fbcfec4d3b2full, 0xe9b5dba58189dbbcull, 0x3956c25bf348b538ull,
            0x59f111f1b605d019ull, 0x923f82a4af194f9bull, 0xab1c5ed5da6d8118ull, 0xd807aa98a3030242ull, 0x12835b0145706fbeull,
            0x243185be4ee4b28cull, 0x550c7dc3d5ffb4e2ull, 0x72be5d74f27b896full, 0x80deb1fe3b1696b1ull, 0x9bdc06a725c71235ull,
            0xc19bf174cf692694ull, 0xe49b69c19ef14ad2ull, 0xefbe4786384f25e3ull, 0x0fc19dc68b8cd5b5ull, 0x240ca1cc77ac9c65ull,
            0x2de92c6f592b0275ull, 0x4a7484aa6ea6e483ull, 0x5cb0a9dcbd41fbd4ull, 0x76f988da831153b5ull, 0x983e5152ee66dfabull,
            0xa831c66d2db43210ull, 0xb00327c898fb213full, 0xbf597fc7beef0ee4ull, 0xc6e00bf33da88fc2ull, 0xd5a79147930aa725ull,
            0x06ca6351e003826full, 0x142929670a0e6e70ull, 0x27b70a8546d22ffcull, 0x2e1b21385c26c926ull, 0x4d2c6dfc5ac42aedull,
            0x53380d139d95b3dfull, 0x650a73548baf63deull, 0x766a0abb3c77b2a8ull, 0x81c2c92e47edaee6ull, 0x92722c851482353bull,
            0xa2bfe8a14cf10364ull, 0xa81a664bbc423001ull, 0xc24b8b70d0f89791ull, 0xc76c51a30654be30ull, 0xd192e819d6ef5218ull,
            0xd69906245565a910ull, 0xf40e35855771202aull, 0x106aa07032bbd1b8ull, 0x19a4c116b8d2d0c8ull, 0x1e376c085141ab53ull,
            0x2748774cdf8eeb99ull, 0x34b0bcb5e19b48a8ull, 0x391c0cb3c5c95a63ull, 0x4ed8aa4ae3418acbull, 0x5b9cca4f7763e373ull,
            0x682e6ff3d6b2b8a3ull, 0x748f82ee5defb2fcull, 0x78a5636f43172f60ull, 0x84c87814a1f0ab72ull, 0x8cc702081a6439ecull,
            0x90befffa23631e28ull, 0xa4506cebde82bde9ull, 0xbef9a3f7b2c67915ull, 0xc67178f2e372532bull, 0xca273eceea26619cull,
            0xd186b8c721c0c207ull, 0xeada7dd6cde0eb1eull, 0xf57d4f7fee6ed178ull, 0x06f067aa72176fbaull, 0x0a637dc5a2c898a6ull,
            0x113f9804bef90daeull, 0x1b710b35131c471bull, 0x28db77f523047d84ull, 0x32caab7b40c72493ull, 0x3c9ebe0a15c9bebcull,
            0x431d67c49c100d4cull, 0x4cc5d4becb3e42b6ull, 0x597f299cfc657e2aull, 0x5fcb6fab3ad6faecull, 0x6c44198c4a475817ull};
        uint64_t w[80];
        for (int i = 0; i < 16; ++i) {
            w[i] = 0;
            for (int j = 0; j < 8; ++j) w[i] = (w[i] << 8) | p[8 * i + j];
        }
        for (int i = 16; i < 80; ++i) {
            const uint64_t s0 = rotr(w[i - 15], 1) ^ rotr(w[i - 15], 8) ^ (w[i - 15] >> 7);
            const uint64_t s1 = rotr(w[i - 2], 19) ^ rotr(w[i - 2], 61) ^ (w[i - 2] >> 6);
            w[i] = w[i - 16] + s0 + w[i - 7] + s1;
        }
        uint64_t a = h[0], b = h[1], c = h[2], d = h[3], e = h[4], f = h[5], g = h[6], hh = h[7];
        for (int i = 0; i < 80; ++i) {
            const uint64_t t1 = hh + (rotr(e, 14) ^ rotr(e, 18) ^ rotr(e, 41)) + ((e & f) ^ (~e & g)) + K[i] + w[i];
            const uint64_t t2 = (rotr(a, 28) ^ rotr(a, 34) ^ rotr(a, 39)) + ((a & b) ^ (a & c) ^ (b & c));
            hh = g; g = f; f = e; e = d + t1; d = c; c = b; b = a; a = t1 + t2;
        }
        h[0] += a; h[1] += b; h[2] += c; h[3] += d; h[4] += e; h[5] += f; h[6] += g; h[7] += hh;
    }
    void digest_short(const uint8_t *data, size_t n, uint8_t out[64]) {  // n <= 111: one block
        uint8_t blk[128] = {0};
        memcpy(blk, data, n);
        blk[n] = 0x80;
        const uint64_t bits = (uint64_t)n * 8;
        for (int k = 0; k < 8; ++k) blk[127 - k] = (uint8_t)(bits >> (8 * k));
        block(blk);
        for (int k = 0; k < 8; ++k)
            for (int j = 0; j < 8; ++j) out[8 * k + j] = (uint8_t)(h[k] >> (56 - 8 * j));
    }
};

struct PyRandom {  // CPython's _random.Random (MT19937)
    uint32_t mt[624];
    int idx = 624;
    void init_genrand(uint32_t s) {
        mt[0] = s;
        for (int i = 1; i < 624; ++i) mt[i] = 1812433253u * (mt[i - 1] ^ (mt[i - 1] >> 30)) + (uint32_t)i;
        idx = 624;
    }
    void init_by_array(const uint32_t *key, size_t len) {
        init_genrand(19650218u);
        size_t i = 1, j = 0;
        for (size_t k = std::max<size_t>(624, len); k; --k) {
            mt[i] = (mt[i] ^ ((mt[i - 1] ^ (mt[i - 1] >> 30)) * 1664525u)) + key[j] + (uint32_t)j;
            if (++i >= 624) {
                mt[0] = mt[623];
                i = 1;
            }
            if (++j >= len) j = 0;
        }
        for (size_t k = 623; k; --k) {
            mt[i] = (mt[i] ^ ((mt[i - 1] ^ (mt[i - 1] >> 30)) * 1566083941u)) - (uint32_t)i;
            if (++i >= 624) {
                mt[0] = mt[623];
                i = 1;
            }
        }
        mt[0] = 0x80000000u;
    }
    // random.seed(a: bytes), version 2: the integer int.from_bytes(a + sha512(a).digest(), "big")
    void seed_bytes(const uint8_t *a, size_t n) {  // n <= 47
        uint8_t big[47 + 64];
        memcpy(big, a, n);
        Sha512().digest_short(a, n, big + n);
        const size_t total = n + 64;
        size_t lead = 0;
        while (lead < total && big[lead] == 0) ++lead;
        const size_t nbytes = total - lead;
        size_t words = std::max<size_t>(1, (nbytes + 3) / 4);
        std::vector<uint32_t> key(words, 0u);
        for (size_t k = 0; k < nbytes; ++k) key[k / 4] |= (uint32_t)big[total - 1 - k] << (8 * (k % 4));
        init_by_array(key.data(), words);
    }
    uint32_t next() {
        if (idx >= 624) {
            static const uint32_t mag01[2] = {0u, 0x9908b0dfu};
            int kk = 0;
            for (; kk < 624 - 397; ++kk) {
                const uint32_t y = (mt[kk] & 0x80000000u) | (mt[kk + 1] & 0x7fffffffu);
                mt[kk] = mt[kk + 397] ^ (y >> 1) ^ mag01[y & 1u];
            }
            for (; kk < 623; ++kk) {
                const uint32_t y = (mt[kk] & 0x80000000u) | (mt[kk + 1] & 0x7fffffffu);
                mt[kk] = mt[kk + (397 - 624)] ^ (y >> 1) ^ mag01[y & 1u];
            }
            const uint32_t y = (mt[623] & 0x80000000u) | (mt[0] & 0x7fffffffu);
            mt[623] = mt[396] ^ (y >> 1) ^ mag01[y & 1u];
            idx = 0;
        }
        uint32_t y = mt[idx++];
        y ^= y >> 11;
        y ^= (y << 7) & 0x9d2c5680u;
        y ^= (y << 15) & 0xefc60000u;
        y ^= y >> 18;
        return y;
    }
    uint32_t randbelow(uint32_t n) {  // Random._randbelow_with_getrandbits, n in [1, 2^31)
        const int k = 32 - __builtin_clz(n);
        uint32_t r = next() >> (32 - k);
        while (r >= n) r = next() >> (32 - k);
        return r;
    }
};

void replace_n(uint8_t *m, int64_t n_rows, int64_t n_cols) {
    if (n_rows <= 0 || n_cols <= 0) return;
    uint8_t digest[32];
    Sha256().digest(m, (uint64_t)(n_rows * n_cols), digest);
    PyRandom rng;
    rng.seed_bytes(digest, 32);
    std::vector<uint8_t> consensus((size_t)n_cols);
    constexpr int BLOCK = 32;
    std::vector<uint32_t> counts((size_t)BLOCK * 256);
    uint8_t order[BLOCK][256];
    int n_order[BLOCK];
    for (int64_t c0 = 0; c0 < n_cols; c0 += BLOCK) {
        const int w = (int)std::min<int64_t>(BLOCK, n_cols - c0);
        std::fill(counts.begin(), counts.end(), 0u);
        for (int b = 0; b < w; ++b) n_order[b] = 0;
        for (int64_t r = 0; r < n_rows; ++r) {
            const uint8_t *row = m + r * n_cols + c0;
            for (int b = 0; b < w; ++b) {
                const uint8_t ch = row[b];
                if (ch == '-' || ch == 'N') continue;
                if (counts[(size_t)b * 256 + ch]++ == 0) order[b][n_order[b]++] = ch;  // Counter keeps first-seen order
            }
        }
        for (int b = 0; b < w; ++b) {
            if (n_order[b] == 0) {
                consensus[(size_t)(c0 + b)] = (uint8_t)"ACGT"[rng.randbelow(4)];
                continue;
            }
            uint32_t top = 0;
            for (int k = 0; k < n_order[b]; ++k) top = std::max(top, counts[(size_t)b * 256 + order[b][k]]);
            uint8_t cand[256];
            uint32_t n_cand = 0;
            for (int k = 0; k < n_order[b]; ++k)
                if (counts[(size_t)b * 256 + order[b][k]] == top) cand[n_cand++] = order[b][k];
            consensus[(size_t)(c0 + b)] = cand[rng.randbelow(n_cand)];
        }
    }
    for (int64_t r = 0; r < n_rows; ++r) {
        uint8_t *row = m + r * n_cols;
        for (int64_t c = 0; c < n_cols; ++c)
            if (row[c] == 'N') row[c] = consensus[(size_t)c];
    }
}

// Mirrors make_prg_b200.utils.io_utils.parse_fasta + the upper-casing of load_alignment_file.
// True when two of the record ids (first whitespace-separated token of each title) are equal.  The common case --
// all distinct -- is settled by sorting 64-bit hashes of the tokens; equal hashes are compared as strings.
bool has_duplicate_ids(const std::string &t, int n_rows) {
    struct Tok { uint64_t h; uint32_t a, n; };
    static thread_local std::vector<Tok> toks;
    toks.clear();
    toks.reserve((size_t)n_rows);
    size_t at = 0;
    auto is_ws = [](unsigned char c) { return c == ' ' || (c >= '\t' && c <= '\r') || (c >= 0x1c && c <= 0x1f); };
    for (int r = 0; r < n_rows; ++r) {
        size_t e = t.find('\n', at);
        if (e == std::string::npos) e = t.size();
        size_t a = at;
        while (a < e && is_ws((unsigned char)t[a])) ++a;
        size_t b = a;
        uint64_t h = 0xcbf29ce484222325ull;
        while (b < e && !is_ws((unsigned char)t[b])) {
            h = (h ^ (unsigned char)t[b]) * 0x100000001b3ull;
            ++b;
        }
        toks.push_back(Tok{h, (uint32_t)a, (uint32_t)(b - a)});
        at = e + 1;
    }
    std::sort(toks.begin(), toks.end(), [](const Tok &x, const Tok &y) { return x.h < y.h; });
    for (size_t i = 0; i + 1 < toks.size(); ++i) {
        for (size_t j = i + 1; j < toks.size() && toks[j].h == toks[i].h; ++j)
            if (toks[j].n == toks[i].n && t.compare(toks[j].a, toks[j].n, t, toks[i].a, toks[i].n) == 0) return true;
    }
    return false;
}

void parse_file(const char *path, ParsedFile &pf, bool avx2, Slab &slab) {
    // the read buffer is reused by the files of one thread (a pool block as the target of read() made the
    // threads of a call run one after the other here: 34 instead of 7.5 ms for 200 files on 8 threads)
    static thread_local std::vector<uint8_t> raw;
    const uint8_t *text = nullptr;
    size_t text_size = 0;
    bool ok;
    if (ends_with(path, ".gz")) {
        ok = read_gzip(path, raw);
        text = raw.data();
        text_size = raw.size();
    } else {
        ok = read_plain(path, nullptr, 0, raw, &text, &text_size);
    }
    if (!ok) {
        pf.status = MPRG_LOAD_IO_ERROR;
        return;
    }
    pf.buf = slab.take(text_size + 32);
    if (!pf.buf) {
        pf.status = MPRG_LOAD_IO_ERROR;
        return;
    }
    const uint8_t *s = text, *end = s + text_size;
    uint8_t *w = pf.buf;
    bool ragged = false;
    int bits = 0;
    int64_t first_len = -1;
    // text before the first title line is skipped
    while (s < end && *s != '>') {
        const uint8_t *le, *nx;
        next_line(s, end, le, nx);
        s = nx;
    }
    while (s < end) {  // *s == '>' : one record per trip
        const uint8_t *le, *nx;
        next_line(s, end, le, nx);
        if (pf.n_rows > 0) pf.titles.push_back('\n');
        {
            // title = line[1:].rstrip() (the ASCII whitespace here; Unicode blanks when the text is decoded)
            const uint8_t *te = le;
            while (te > s + 1 && (te[-1] == ' ' || is_trailing_ws(te[-1]))) --te;
            pf.titles.append((const char *)s + 1, (size_t)(te - s - 1));
        }
        s = nx;
        uint8_t *row_start = w;
        if (s < end && *s != '>')
            s = avx2 ? copy_record_avx2(s, end, w, bits, row_start) : copy_record_scalar(s, end, w, bits, row_start);
        const int64_t len = w - row_start;
        if (first_len < 0)
            first_len = len;
        else if (len != first_len)
            ragged = true;
        pf.n_rows++;
    }
    slab.give_back(pf.buf, text_size + 32, (size_t)(w - pf.buf) + 32);
    if (pf.n_rows == 0) {
        pf.status = MPRG_LOAD_NO_RECORDS;
        return;
    }
    // titles may hold any valid UTF-8 (Biopython reads text); a title that does not decode makes the
    // reference's read fail, so that file goes back to the Python loader for the exception
    if (!valid_utf8((const uint8_t *)pf.titles.data(), pf.titles.size())) bits |= 1;
    if (bits & 1) {
        pf.status = MPRG_LOAD_NOT_ASCII;
        return;
    }
    if (ragged) {
        pf.status = MPRG_LOAD_RAGGED;
        return;
    }
    if (has_duplicate_ids(pf.titles, pf.n_rows)) pf.flags |= MPRG_LOAD_FLAG_DUPLICATE_IDS;
    pf.n_cols = (int32_t)first_len;
    pf.matrix_bytes = (int64_t)pf.n_rows * first_len;
    if (first_len > INT32_MAX) pf.status = MPRG_LOAD_IO_ERROR;
    if ((bits & 2) && pf.status == MPRG_LOAD_OK) {
        pf.flags |= MPRG_LOAD_FLAG_HAS_N;
        replace_n(pf.buf, pf.n_rows, pf.n_cols);
    }
}


// ------------------------------------------------------------------------------------------------
// ASCII rows -> the 4-bit device layout, on the host (what pack_rows_kernel of batch.cu produces): rows
// padded to 16 bytes = 32 columns with MPRG_SYM_PAD, column c of a chunk in nibble c / 4 of word c % 4.
// The loader emits its matrices this way so that half the bytes cross PCIe and no pack kernel runs
// (mprg_build_packed).  flags: bit0 a character outside the alphabet, bit1 N, bit2 RYKMSW, bit3 an even
// symbol code (M S W N) -- exactly the flags of the pack kernel.
// ------------------------------------------------------------------------------------------------
struct PackTables {
    uint8_t code[256], cls[16];
    PackTables() {
        memset(code, MPRG_SYM_PAD, sizeof(code));
        const char *alphabet = MPRG_ALPHABET;
        for (int i = 0; i < 16; ++i) {
            const char ch = alphabet[i];
            if (ch == '?') continue;
            code[(uint8_t)ch] = (uint8_t)i;
            if (ch >= 'A' && ch <= 'Z') code[(uint8_t)(ch - 'A' + 'a')] = (uint8_t)i;
        }
        for (int c = 0; c < 16; ++c) {
            int f = 0;
            if (c == MPRG_SYM_PAD) f |= 1;
            else if (c == MPRG_SYM_N) f |= 2;
            else if (c != MPRG_SYM_GAP && (c & 9) != 1) f |= 4;
            if (c != MPRG_SYM_GAP && !(c & 1)) f |= 8;
            cls[c] = (uint8_t)f;
        }
    }
};
const PackTables &pack_tables() {
    static const PackTables t;
    return t;
}

// one chunk (up to 32 valid columns at src) -> 16 bytes
inline int pack_chunk_scalar(const uint8_t *src, int valid, uint8_t *dst) {
    const PackTables &t = pack_tables();
    uint32_t words[4] = {0, 0, 0, 0};
    int f = 0;
    for (int c = 0; c < 32; ++c) {
        uint32_t code = MPRG_SYM_PAD;
        if (c < valid) {
            code = t.code[src[c]];
            f |= t.cls[code];
        }
        words[c & 3] |= code << (4 * (c >> 2));
    }
    memcpy(dst, words, 16);
    return f;
}

__attribute__((target("avx2"))) int pack_row_avx2(const uint8_t *src, int cols, uint8_t *dst) {
    // letters by their low five bits: two 16-entry tables ('@'..'O', 'P'..'_'), made once (this runs per row)
    struct Vec {
        alignas(32) uint8_t lo[32], hi[32], cls[32];
        Vec() {
            const PackTables &t = pack_tables();
            for (int k = 0; k < 16; ++k) {
                lo[k] = lo[k + 16] = t.code[0x40 + k];
                hi[k] = hi[k + 16] = t.code[0x50 + k];
                cls[k] = cls[k + 16] = t.cls[k];
            }
            lo[0] = lo[16] = MPRG_SYM_PAD;  // '@' / '`' are no letters
        }
    };
    static const Vec tv;
    const __m256i t_lo = _mm256_load_si256((const __m256i *)tv.lo), t_hi = _mm256_load_si256((const __m256i *)tv.hi);
    const __m256i t_cls = _mm256_load_si256((const __m256i *)tv.cls);
    const __m256i v_dash = _mm256_set1_epi8('-'), v_case = _mm256_set1_epi8(0x20), v_a = _mm256_set1_epi8('a' - 1);
    const __m256i v_z = _mm256_set1_epi8('z' + 1), v_0f = _mm256_set1_epi8(0x0f), v_10 = _mm256_set1_epi8(0x10);
    const __m256i v_pad = _mm256_set1_epi8(MPRG_SYM_PAD);
    const __m256i shuf = _mm256_setr_epi8(0, 4, 8, 12, 1, 5, 9, 13, 2, 6, 10, 14, 3, 7, 11, 15, 0, 4, 8, 12, 1, 5, 9, 13, 2, 6,
                                          10, 14, 3, 7, 11, 15);
    const __m256i m1 = _mm256_set1_epi16(0x1001), m2 = _mm256_set1_epi32(0x01000001);
    __m256i acc = _mm256_setzero_si256();
    int c = 0;
    for (; c + 32 <= cols; c += 32, dst += 16) {
        const __m256i v = _mm256_loadu_si256((const __m256i *)(src + c));
        const __m256i low = _mm256_or_si256(v, v_case);
        const __m256i letter = _mm256_and_si256(_mm256_cmpgt_epi8(low, v_a), _mm256_cmpgt_epi8(v_z, low));
        const __m256i idx = _mm256_and_si256(v, v_0f);
        const __m256i upper_half = _mm256_cmpeq_epi8(_mm256_and_si256(v, v_10), v_10);
        __m256i code = _mm256_blendv_epi8(_mm256_shuffle_epi8(t_lo, idx), _mm256_shuffle_epi8(t_hi, idx), upper_half);
        code = _mm256_blendv_epi8(v_pad, code, letter);
        code = _mm256_andnot_si256(_mm256_cmpeq_epi8(v, v_dash), code);  // '-' = 0
        acc = _mm256_or_si256(acc, _mm256_shuffle_epi8(t_cls, code));
        const __m256i x = _mm256_shuffle_epi8(code, shuf);
        const __m256i q = _mm256_madd_epi16(_mm256_maddubs_epi16(x, m1), m2);
        const __m128i out = _mm_or_si128(_mm256_castsi256_si128(q), _mm_slli_epi32(_mm256_extracti128_si256(q, 1), 16));
        _mm_storeu_si128((__m128i *)dst, out);
    }
    alignas(32) uint8_t accb[32];
    _mm256_store_si256((__m256i *)accb, acc);
    int f = 0;
    for (int k = 0; k < 32; ++k) f |= accb[k];
    if (c < cols) f |= pack_chunk_scalar(src + c, cols - c, dst);
    return f;
}

int pack_row_scalar(const uint8_t *src, int cols, uint8_t *dst) {
    int f = 0;
    for (int c = 0; c < cols; c += 32, dst += 16) f |= pack_chunk_scalar(src + c, std::min(32, cols - c), dst);
    return f;
}

inline int64_t packed_stride(int32_t cols) { return ((int64_t)cols + 31) / 32 * 16; }

int pack_matrix(const uint8_t *ascii, int32_t rows, int32_t cols, uint8_t *out, bool avx2) {
    const int64_t stride = packed_stride(cols);
    int f = 0;
    for (int32_t r = 0; r < rows; ++r) {
        const uint8_t *src = ascii + (int64_t)r * cols;
        uint8_t *dst = out + (int64_t)r * stride;
        f |= avx2 ? pack_row_avx2(src, cols, dst) : pack_row_scalar(src, cols, dst);
    }
    return f;
}
}  // namespace

extern "C" int mprg_pack_rows(const uint8_t *h_ascii, int32_t n_rows, int32_t n_cols, uint8_t *h_packed,
                              int64_t capacity, int32_t *flags) {
    if (n_rows < 0 || n_cols < 0 || (!h_ascii && (int64_t)n_rows * n_cols > 0) || !h_packed) return MPRG_E_BAD_ARG;
    if (capacity < packed_stride(n_cols) * n_rows) return MPRG_E_BAD_ARG;
    const bool avx2 = __builtin_cpu_supports("avx2") && !getenv("MPRG_NO_AVX2");
    const int f = pack_matrix(h_ascii, n_rows, n_cols, h_packed, avx2);
    if (flags) *flags = f;
    return MPRG_OK;
}

extern "C" int mprg_replace_n(uint8_t *h_ascii, int32_t n_rows, int32_t n_cols) {
    if (!h_ascii || n_rows < 0 || n_cols < 0) return MPRG_E_BAD_ARG;
    replace_n(h_ascii, n_rows, n_cols);
    return MPRG_OK;
}

extern "C" void mprg_fasta_free(mprg_msa_set *set);

struct mprg_msa_set {
    int32_t n = 0;
    uint8_t *ascii = nullptr;
    int64_t ascii_bytes = 0;
    bool pinned = false;
    size_t capacity = 0;
    std::vector<int64_t> offsets;
    std::vector<int32_t> n_rows, n_cols, status, flags;
    std::vector<std::string> titles;
    // packed mode (MPRG_LOADMODE_PACKED): the matrices in the 4-bit device layout, loci back to back
    uint8_t *packed = nullptr;
    int64_t packed_bytes = 0;
    bool packed_pinned = false;
    size_t packed_capacity = 0;
    std::vector<int64_t> packed_offsets;
    std::vector<int32_t> alphabet_flags;  // per locus: the flags of the pack kernel (mprg_batch_flags)
};

extern "C" int mprg_fasta_load(const char *const *paths, int32_t n_files, int32_t n_threads, int32_t mode,
                               mprg_msa_set **out) {
    if (!out || n_files < 0 || (n_files > 0 && !paths)) return MPRG_E_BAD_ARG;
    const int pin = mode & MPRG_LOADMODE_PIN;
    const bool want_packed = (mode & MPRG_LOADMODE_PACKED) != 0;
    const bool want_ascii = !want_packed || (mode & MPRG_LOADMODE_KEEP_ASCII) != 0;
    *out = nullptr;
    std::vector<ParsedFile> files((size_t)n_files);
    const bool avx2 = __builtin_cpu_supports("avx2") && !getenv("MPRG_NO_AVX2");
    n_threads = std::max(1, std::min(n_threads, std::max(n_files, 1)));
    std::vector<Slab> slabs((size_t)n_threads);
    parallel_for_t(n_files, n_threads, [&](int i, int t) { parse_file(paths[i], files[(size_t)i], avx2, slabs[(size_t)t]); });
    mprg_msa_set *set = new mprg_msa_set();
    set->n = n_files;
    set->offsets.resize((size_t)n_files);
    set->n_rows.resize((size_t)n_files);
    set->n_cols.resize((size_t)n_files);
    set->status.resize((size_t)n_files);
    set->flags.resize((size_t)n_files);
    set->titles.resize((size_t)n_files);
    int64_t total = 0;
    for (int i = 0; i < n_files; ++i) {
        ParsedFile &pf = files[(size_t)i];
        const bool good = pf.status == MPRG_LOAD_OK;
        set->offsets[(size_t)i] = total;
        set->n_rows[(size_t)i] = good ? pf.n_rows : 0;
        set->n_cols[(size_t)i] = good ? pf.n_cols : 0;
        set->status[(size_t)i] = pf.status;
        set->flags[(size_t)i] = pf.flags;
        set->titles[(size_t)i].swap(pf.titles);
        if (good) total += pf.matrix_bytes;
    }
    set->ascii_bytes = want_ascii ? total : 0;
    if (want_ascii) {
        const size_t alloc = (size_t)std::max<int64_t>(total, 1);
        // in packed mode the text is only kept for callers that want the rows back: ordinary memory
        if (pin && !want_packed) set->ascii = pinned_pool().acquire(alloc, set->capacity);
        if (set->ascii) {
            set->pinned = true;
        } else {  // without a device the buffer is ordinary memory
            set->ascii = huge_alloc(alloc);
            if (!set->ascii) {
                delete set;
                return MPRG_E_INTERNAL;
            }
        }
    }
    if (want_packed) {
        set->packed_offsets.resize((size_t)n_files);
        set->alphabet_flags.assign((size_t)n_files, 0);
        int64_t ptotal = 0;
        for (int i = 0; i < n_files; ++i) {
            set->packed_offsets[(size_t)i] = ptotal;
            ptotal += packed_stride(set->n_cols[(size_t)i]) * set->n_rows[(size_t)i];
        }
        set->packed_bytes = ptotal;
        const size_t palloc = (size_t)std::max<int64_t>(ptotal, 1);
        if (pin) set->packed = pinned_pool().acquire(palloc, set->packed_capacity);
        if (set->packed) {
            set->packed_pinned = true;
        } else {
            set->packed = huge_alloc(palloc);
            if (!set->packed) {
                mprg_fasta_free(set);
                return MPRG_E_INTERNAL;
            }
        }
    }
    parallel_for(n_files, n_threads, [&](int i) {
        ParsedFile &pf = files[(size_t)i];
        if (pf.status != MPRG_LOAD_OK || !pf.matrix_bytes) return;
        if (want_packed)
            set->alphabet_flags[(size_t)i] =
                pack_matrix(pf.buf, pf.n_rows, pf.n_cols, set->packed + set->packed_offsets[(size_t)i], avx2);
        if (want_ascii) memcpy(set->ascii + set->offsets[(size_t)i], pf.buf, (size_t)pf.matrix_bytes);
    });
    *out = set;
    return MPRG_OK;
}

extern "C" int mprg_fasta_packed(const mprg_msa_set *set, uint8_t **h_packed, int64_t *packed_bytes,
                                 const int64_t **h_packed_offsets, const int32_t **alphabet_flags) {
    if (!set || !set->packed) return MPRG_E_BAD_ARG;
    if (h_packed) *h_packed = set->packed;
    if (packed_bytes) *packed_bytes = set->packed_bytes;
    if (h_packed_offsets) *h_packed_offsets = set->packed_offsets.data();
    if (alphabet_flags) *alphabet_flags = set->alphabet_flags.data();
    return MPRG_OK;
}

extern "C" void mprg_fasta_free(mprg_msa_set *set) {
    if (!set) return;
    if (set->ascii) {
        if (set->pinned)
            pinned_pool().release(set->ascii, set->capacity);
        else
            free(set->ascii);
    }
    if (set->packed) {
        if (set->packed_pinned)
            pinned_pool().release(set->packed, set->packed_capacity);
        else
            free(set->packed);
    }
    delete set;
}

extern "C" int mprg_fasta_info(const mprg_msa_set *set, int32_t *n_loci, uint8_t **h_ascii, int64_t *ascii_bytes,
                               const int64_t **h_offsets, const int32_t **n_rows, const int32_t **n_cols,
                               const int32_t **status, const int32_t **flags) {
    if (!set) return MPRG_E_BAD_ARG;
    if (n_loci) *n_loci = set->n;
    if (h_ascii) *h_ascii = set->ascii;
    if (ascii_bytes) *ascii_bytes = set->ascii_bytes;
    if (h_offsets) *h_offsets = set->offsets.data();
    if (n_rows) *n_rows = set->n_rows.data();
    if (n_cols) *n_cols = set->n_cols.data();
    if (status) *status = set->status.data();
    if (flags) *flags = set->flags.data();
    return MPRG_OK;
}

extern "C" const char *mprg_fasta_titles(const mprg_msa_set *set, int32_t locus, int64_t *length) {
    if (!set || locus < 0 || locus >= set->n) return nullptr;
    if (length) *length = (int64_t)set->titles[(size_t)locus].size();
    return set->titles[(size_t)locus].data();
}

// ------------------------------------------------------------------------------------------------
// writers
// ------------------------------------------------------------------------------------------------
namespace {

// output buffer of an encoder: allocated to an upper bound WITHOUT being zero-filled (the tail that is
// never written is never touched, so never paged in), then cut to the length written
template <typename T>
struct RawBuf {
    std::unique_ptr<T[]> p;
    size_t n = 0;
    void allocate(size_t capacity) {
        p.reset(new T[capacity > 0 ? capacity : 1]);
        n = 0;
    }
    T *data() { return p.get(); }
    const T *data() const { return p.get(); }
    size_t size() const { return n; }
    bool empty() const { return n == 0; }
};

const char GFA_HEADER[] = "H\tVN:Z:1.0\tbn:Z:--linear --singlearr\n";

inline bool is_py_space(unsigned char c) {  // str.split() separators within ASCII
    return c == ' ' || (c >= 9 && c <= 13) || (c >= 0x1c && c <= 0x1f);
}

// character classes of a PRG string: 1..4 = A C G T (either case), 5 = digit, 6 = separator, 0 = other
struct PrgClasses {
    uint8_t of[256];
    PrgClasses() {
        memset(of, 0, sizeof(of));
        const char *bases = "ACGT";
        for (int k = 0; k < 4; ++k) {
            of[(unsigned char)bases[k]] = (uint8_t)(k + 1);
            of[(unsigned char)(bases[k] + 32)] = (uint8_t)(k + 1);
        }
        for (int c = '0'; c <= '9'; ++c) of[c] = 5;
        for (int c = 0; c < 256; ++c)
            if (is_py_space((unsigned char)c)) of[c] = 6;
    }
};
const PrgClasses PRG_CLASSES;

// prg_encoder.py:44-91.  Returns 0, or MPRG_ENC_* on the reference's exceptions.
int encode_prg(const char *prg, int64_t len, RawBuf<uint32_t> &out) {
    out.allocate((size_t)len);  // one value per base or marker: never more than characters
    uint32_t *o = out.data();
    const uint8_t *cls = PRG_CLASSES.of;
    const unsigned char *p = reinterpret_cast<const unsigned char *>(prg);
    // odd marker -> times seen; sites close in LIFO order, so the search runs from the back
    std::vector<std::pair<uint64_t, int>> seen;
    int64_t i = 0;
    while (i < len) {
        if (cls[p[i]] == 6) {
            ++i;
            continue;
        }
        const int64_t u0 = i;
        uint32_t *o0 = o;
        unsigned all = 0xff, any = 0;
        while (i < len) {  // written as bases while scanning; rewound if the unit is something else
            const unsigned c = cls[p[i]];
            if (c == 6) break;
            *o++ = c;
            all &= (c >= 1 && c <= 4) ? 1u : (c == 5 ? 2u : 0u);
            any = 1;
            ++i;
        }
        (void)any;
        if (all & 1u) continue;  // a unit of bases
        o = o0;
        if (!(all & 2u)) return MPRG_ENC_INVALID_UNIT;
        uint64_t m = 0;
        for (int64_t j = u0; j < i; ++j) {
            m = m * 10 + (uint64_t)(p[j] - '0');
            if (m > 0xFFFFFFFFull) return MPRG_ENC_OVERFLOW;
        }
        if ((m & 1) == 0) {
            *o++ = (uint32_t)m;
            continue;
        }
        int times = 0;
        for (size_t k = seen.size(); k-- > 0;)
            if (seen[k].first == m) {
                times = ++seen[k].second;
                break;
            }
        if (times == 0) {
            seen.emplace_back(m, 1);
            times = 1;
        }
        if (times > 2) return MPRG_ENC_ODD_MARKER_REPEATED;
        if (times == 2 && m + 1 > 0xFFFFFFFFull) return MPRG_ENC_OVERFLOW;
        *o++ = (uint32_t)(times == 1 ? m : m + 1);
    }
    out.n = (size_t)(o - out.data());
    return 0;
}

// gfa.py:39-109 as one recursive descent over the token stream (markers are " <digits> ", anything
// else between spaces is sequence); segment / link ids come out in the reference's order.  The text is
// written through a raw pointer into a buffer sized to its upper bound.
struct GfaBuilder {
    const char *p;
    int64_t len, i = 0;
    char *w = nullptr;
    uint64_t gfa_id = 0;
    bool bad = false;
    std::vector<uint64_t> ids;  // ends of the alleles of the open sites (a stack shared by the recursion)

    enum Tok { END, LITERAL, MARKER };
    Tok tok = END;
    int64_t lit0 = 0, lit1 = 0;
    uint64_t marker = 0;

    void advance() {
        while (i < len) {
            if (p[i] == ' ') {
                int64_t j = i + 1;
                uint64_t m = 0;
                while (j < len && p[j] >= '0' && p[j] <= '9') {
                    m = m * 10 + (uint64_t)(p[j] - '0');
                    ++j;
                }
                if (j > i + 1 && j < len && p[j] == ' ') {
                    tok = MARKER;
                    marker = m;
                    i = j + 1;
                    return;
                }
                ++i;  // a space that is not part of a marker is skipped
                continue;
            }
            lit0 = i;
            const void *sp = memchr(p + i, ' ', (size_t)(len - i));
            i = sp ? (int64_t)((const char *)sp - p) : len;
            lit1 = i;
            tok = LITERAL;
            return;
        }
        tok = END;
    }
    inline void put(const char *s, size_t n) {
        memcpy(w, s, n);
        w += n;
    }
    inline void put_uint(uint64_t v) {
        char tmp[24];
        int n = 0;
        do {
            tmp[n++] = (char)('0' + v % 10);
            v /= 10;
        } while (v);
        while (n) *w++ = tmp[--n];
    }
    void segment(int64_t a, int64_t b, const std::string *extra) {
        put("S\t", 2);
        put_uint(gfa_id);
        *w++ = '\t';
        if (extra && !extra->empty())
            put(extra->data(), extra->size());
        else if (b > a)
            put(p + a, (size_t)(b - a));
        else
            *w++ = '*';
        put("\tRC:i:0\n", 8);
    }
    void link(uint64_t a, uint64_t b) {
        put("L\t", 2);
        put_uint(a);
        put("\t+\t", 3);
        put_uint(b);
        put("\t+\t0M\n", 6);
    }
    // One (sub)string of the PRG: literal? (site literal?)*.  open_marker = odd marker of the enclosing
    // site (0 at top level).  Returns the id of the segment that ends it; on return `tok` is the token
    // that ended the sequence (END, the closing odd marker, or the even separator).
    uint64_t sequence(uint64_t open_marker, int depth) {
        const size_t base = ids.size();  // ids[base..) = ends of the alleles of the site just closed
        int64_t a = 0, b = 0;            // pending literal as one slice of the input ...
        std::string joined;              // ... or, when several literal tokens follow each other, their join
        while (!bad) {
            if (tok == LITERAL) {
                if (b > a || !joined.empty()) {
                    if (joined.empty()) joined.assign(p + a, (size_t)(b - a));
                    joined.append(p + lit0, (size_t)(lit1 - lit0));
                } else {
                    a = lit0;
                    b = lit1;
                }
                advance();
                continue;
            }
            if (tok == MARKER && (marker & 1) && marker != open_marker) {
                if (depth > 512) {
                    bad = true;
                    break;
                }
                const uint64_t site = marker;
                segment(a, b, &joined);
                const uint64_t pre = gfa_id++;
                for (size_t k = base; k < ids.size(); ++k) link(ids[k], pre);
                ids.resize(base);
                int alleles = 0;
                advance();
                while (!bad) {
                    link(pre, gfa_id);
                    const uint64_t last = sequence(site, depth + 1);
                    ids.push_back(last);
                    ++alleles;
                    if (tok == MARKER && marker == site + 1) {  // next allele
                        advance();
                        continue;
                    }
                    if (tok == MARKER && marker == site) {  // site closed
                        advance();
                        break;
                    }
                    bad = true;  // unterminated site or a foreign even marker
                }
                if (alleles < 2) bad = true;
                a = b = 0;
                joined.clear();
                continue;
            }
            if (tok == MARKER && !(marker & 1) && marker != open_marker + 1) bad = true;
            if (tok == MARKER && open_marker == 0) bad = true;
            break;  // END, my closing marker or my separator
        }
        segment(a, b, &joined);
        for (size_t k = base; k < ids.size(); ++k) link(ids[k], gfa_id);
        ids.resize(base);
        return gfa_id++;
    }
};

int prg_to_gfa(const char *prg, int64_t len, RawBuf<char> &out) {
    GfaBuilder g;
    g.p = prg;
    g.len = len;
    // upper bound of the text: every input character once, and per marker token at most two segment
    // lines and three link lines (each id < 2 * markers + 2, so at most 20 digits)
    size_t spaces = 0;
    for (int64_t k = 0; k < len; ++k) spaces += prg[k] == ' ';
    const size_t tokens = spaces / 2 + 2;
    out.allocate(sizeof(GFA_HEADER) + (size_t)len + tokens * (2 * 36 + 3 * 56) + 64);
    g.w = out.data();
    g.put(GFA_HEADER, sizeof(GFA_HEADER) - 1);
    g.advance();
    g.sequence(0, 0);
    if (g.bad || g.tok != GfaBuilder::END) return MPRG_ENC_INVALID_UNIT;
    out.n = (size_t)(g.w - out.data());
    return 0;
}

// ---- stored zip archive (what zipfile.ZipFile(path, "w") produces: ZIP_STORED) -------------------
struct ZipEntry {
    std::string name;
    uint32_t crc, size;
    uint64_t offset;
};

// Output files are written under a temporary name next to their final one and renamed when complete, so
// that an aborted run leaves no truncated archive behind (and nothing a rerun without -F would trip over).
std::string temp_name_of(const std::string &path) { return path + ".tmp" + std::to_string((long long)getpid()); }

struct ZipFile {
    FILE *f = nullptr;
    uint64_t pos = 0;
    std::vector<ZipEntry> entries;
    uint16_t dos_time = 0, dos_date = 0;
    std::string final_path, tmp_path;
    void discard() {
        if (f) fclose(f);
        f = nullptr;
        if (!tmp_path.empty()) unlink(tmp_path.c_str());
        tmp_path.clear();
    }

    static void le16(std::string &s, uint16_t v) {
        s.push_back((char)(v & 0xff));
        s.push_back((char)(v >> 8));
    }
    static void le32(std::string &s, uint32_t v) {
        for (int k = 0; k < 4; ++k) s.push_back((char)((v >> (8 * k)) & 0xff));
    }
    static void le64(std::string &s, uint64_t v) {
        for (int k = 0; k < 8; ++k) s.push_back((char)((v >> (8 * k)) & 0xff));
    }
    bool open_path(const std::string &path) {
        final_path = path;
        tmp_path = temp_name_of(path);
        f = fopen(tmp_path.c_str(), "wb");
        if (!f) {
            tmp_path.clear();
            return false;
        }
        setvbuf(f, nullptr, _IOFBF, 1 << 20);
        time_t now = time(nullptr);
        struct tm tmv;
        localtime_r(&now, &tmv);
        const int year = std::max(tmv.tm_year + 1900, 1980);
        dos_date = (uint16_t)(((year - 1980) << 9) | ((tmv.tm_mon + 1) << 5) | tmv.tm_mday);
        dos_time = (uint16_t)((tmv.tm_hour << 11) | (tmv.tm_min << 5) | (tmv.tm_sec / 2));
        return true;
    }
    bool put(const std::string &s) {
        if (fwrite(s.data(), 1, s.size(), f) != s.size()) return false;
        pos += s.size();
        return true;
    }
    bool add(const std::string &name, const void *data, size_t n, uint32_t crc) {
        if (n > 0xFFFFFFFEull || name.size() > 0xFFFF) return false;
        std::string h;
        le32(h, 0x04034b50u);
        le16(h, 20);
        le16(h, 0x0800);  // names are UTF-8
        le16(h, 0);       // stored
        le16(h, dos_time);
        le16(h, dos_date);
        le32(h, crc);
        le32(h, (uint32_t)n);
        le32(h, (uint32_t)n);
        le16(h, (uint16_t)name.size());
        le16(h, 0);
        h += name;
        entries.push_back({name, crc, (uint32_t)n, pos});
        if (!put(h)) return false;
        if (n && fwrite(data, 1, n, f) != n) return false;
        pos += n;
        return true;
    }
    bool finish() {
        const uint64_t cd_offset = pos;
        std::string cd;
        for (const ZipEntry &e : entries) {
            const bool big = e.offset >= 0xFFFFFFFFull;
            le32(cd, 0x02014b50u);
            le16(cd, (uint16_t)((3 << 8) | (big ? 45 : 20)));
            le16(cd, big ? 45 : 20);
            le16(cd, 0x0800);
            le16(cd, 0);
            le16(cd, dos_time);
            le16(cd, dos_date);
            le32(cd, e.crc);
            le32(cd, e.size);
            le32(cd, e.size);
            le16(cd, (uint16_t)e.name.size());
            le16(cd, big ? 12 : 0);
            le16(cd, 0);
            le16(cd, 0);
            le16(cd, 0);
            le32(cd, 0600u << 16);
            le32(cd, big ? 0xFFFFFFFFu : (uint32_t)e.offset);
            cd += e.name;
            if (big) {
                le16(cd, 1);
                le16(cd, 8);
                le64(cd, e.offset);
            }
        }
        const uint64_t cd_size = cd.size(), n = entries.size();
        if (n > 0xFFFE || cd_offset >= 0xFFFFFFFFull || cd_size >= 0xFFFFFFFFull) {
            le32(cd, 0x06064b50u);  // zip64 end of central directory
            le64(cd, 44);
            le16(cd, 45);
            le16(cd, 45);
            le32(cd, 0);
            le32(cd, 0);
            le64(cd, n);
            le64(cd, n);
            le64(cd, cd_size);
            le64(cd, cd_offset);
            le32(cd, 0x07064b50u);  // locator
            le32(cd, 0);
            le64(cd, cd_offset + cd_size);
            le32(cd, 1);
        }
        le32(cd, 0x06054b50u);
        le16(cd, 0);
        le16(cd, 0);
        le16(cd, (uint16_t)std::min<uint64_t>(n, 0xFFFF));
        le16(cd, (uint16_t)std::min<uint64_t>(n, 0xFFFF));
        le32(cd, (uint32_t)std::min<uint64_t>(cd_size, 0xFFFFFFFFull));
        le32(cd, (uint32_t)std::min<uint64_t>(cd_offset, 0xFFFFFFFFull));
        le16(cd, 0);
        bool ok = put(cd);
        ok = (fclose(f) == 0) && ok;
        f = nullptr;
        if (ok) ok = rename(tmp_path.c_str(), final_path.c_str()) == 0;
        if (!ok) unlink(tmp_path.c_str());
        tmp_path.clear();
        return ok;
    }
    // appends the members of a finished stored archive written by this class (a shard's part): its data
    // region is copied as one block, the directory entries move by the current position
    bool append_archive(const std::string &part_path, std::string &err);
};

bool write_whole(const std::string &path, const void *data, size_t n) {
    const std::string tmp = temp_name_of(path);
    FILE *f = fopen(tmp.c_str(), "wb");
    if (!f) return false;
    bool ok = n == 0 || fwrite(data, 1, n, f) == n;
    ok = (fclose(f) == 0) && ok;
    if (ok) ok = rename(tmp.c_str(), path.c_str()) == 0;
    if (!ok) unlink(tmp.c_str());
    return ok;
}

bool read_whole(const std::string &path, std::string &out) {
    FILE *f = fopen(path.c_str(), "rb");
    if (!f) return false;
    out.clear();
    char buf[1 << 16];
    size_t n;
    while ((n = fread(buf, 1, sizeof(buf), f)) > 0) out.append(buf, n);
    const bool ok = !ferror(f);
    fclose(f);
    return ok;
}

inline uint16_t rd16(const uint8_t *p) { return (uint16_t)(p[0] | (p[1] << 8)); }
inline uint32_t rd32(const uint8_t *p) { return (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24); }
inline uint64_t rd64(const uint8_t *p) { return (uint64_t)rd32(p) | ((uint64_t)rd32(p + 4) << 32); }

// Central directory of a stored archive held in memory: entries with their local-header offsets and the
// offset where the directory starts (= size of the data region).  False when the bytes are not such a file.
bool parse_stored_zip(const std::string &z, std::vector<ZipEntry> &entries, uint64_t &cd_offset) {
    entries.clear();
    if (z.size() < 22) return false;
    const uint8_t *b = (const uint8_t *)z.data();
    const size_t eocd = z.size() - 22;  // archives written here carry no comment
    if (rd32(b + eocd) != 0x06054b50u) return false;
    uint64_t n = rd16(b + eocd + 10), cd_size = rd32(b + eocd + 12);
    cd_offset = rd32(b + eocd + 16);
    if (n == 0xFFFF || cd_size == 0xFFFFFFFFull || cd_offset == 0xFFFFFFFFull) {
        if (eocd < 20 + 56 || rd32(b + eocd - 20) != 0x07064b50u) return false;
        const uint64_t z64 = rd64(b + eocd - 20 + 8);
        if (z64 + 56 > z.size() || rd32(b + z64) != 0x06064b50u) return false;
        n = rd64(b + z64 + 32);
        cd_size = rd64(b + z64 + 40);
        cd_offset = rd64(b + z64 + 48);
    }
    if (cd_offset + cd_size > z.size()) return false;
    size_t p = (size_t)cd_offset;
    entries.reserve((size_t)n);
    for (uint64_t i = 0; i < n; ++i) {
        if (p + 46 > z.size() || rd32(b + p) != 0x02014b50u) return false;
        if (rd16(b + p + 10) != 0) return false;  // stored members only
        ZipEntry e;
        e.crc = rd32(b + p + 16);
        e.size = rd32(b + p + 24);
        const size_t nlen = rd16(b + p + 28), xlen = rd16(b + p + 30), clen = rd16(b + p + 32);
        e.offset = rd32(b + p + 42);
        if (p + 46 + nlen + xlen + clen > z.size()) return false;
        e.name.assign(z.data() + p + 46, nlen);
        if (e.offset == 0xFFFFFFFFull) {
            size_t x = p + 46 + nlen;
            const size_t xe = x + xlen;
            bool found = false;
            while (x + 4 <= xe) {
                const uint16_t id = rd16(b + x), sz = rd16(b + x + 2);
                if (id == 1 && sz >= 8) {
                    e.offset = rd64(b + x + 4);
                    found = true;
                    break;
                }
                x += 4 + sz;
            }
            if (!found) return false;
        }
        entries.push_back(std::move(e));
        p += 46 + nlen + xlen + clen;
    }
    return true;
}

bool ZipFile::append_archive(const std::string &part_path, std::string &err) {
    std::string z;
    std::vector<ZipEntry> part;
    uint64_t cd_offset = 0;
    if (!read_whole(part_path, z) || !parse_stored_zip(z, part, cd_offset)) {
        err = "cannot read the archive part " + part_path;
        return false;
    }
    if (cd_offset && fwrite(z.data(), 1, (size_t)cd_offset, f) != cd_offset) {
        err = "cannot write " + final_path + ": " + strerror(errno);
        return false;
    }
    for (ZipEntry &e : part) {
        e.offset += pos;
        entries.push_back(std::move(e));
    }
    pos += cd_offset;
    return true;
}

// data of one member of a stored archive held in memory
bool member_data(const std::string &z, const ZipEntry &e, const char *&data, size_t &n) {
    const uint8_t *b = (const uint8_t *)z.data();
    if (e.offset + 30 > z.size() || rd32(b + e.offset) != 0x04034b50u) return false;
    const size_t start = (size_t)e.offset + 30 + rd16(b + e.offset + 26) + rd16(b + e.offset + 28);
    if (start + e.size > z.size()) return false;
    data = z.data() + start;
    n = e.size;
    return true;
}

struct Encoded {
    std::string name;
    RawBuf<uint32_t> bin;
    RawBuf<char> gfa;
    uint32_t crc_bin = 0, crc_gfa = 0;
    int rc = 0;
};

}  // namespace

struct mprg_writer {
    std::string prefix, err;
    int what = 0;
    std::vector<std::pair<std::string, std::string>> fa;  // (name, prg) for the sorted .prg.fa
    Encoded first;  // kept back until a second locus shows up: one locus => plain files, no archive
    int64_t n_added = 0, bytes = 0;
    ZipFile zbin, zgfa;
    bool zips_open = false;
    ZipFile zds;  // <prefix>.update_DS.zip: one table-shaped member per locus (mprg_writer_add_ds)
    bool ds_open = false;
    int64_t n_ds = 0;
};

extern "C" int mprg_encode_prg(const char *prg, int64_t length, uint32_t *out, int64_t capacity, int64_t *n) {
    if ((!prg && length > 0) || length < 0 || !n) return MPRG_E_BAD_ARG;
    RawBuf<uint32_t> v;
    const int rc = encode_prg(prg, length, v);
    if (rc) return rc;
    *n = (int64_t)v.size();
    if (out && capacity >= (int64_t)v.size() && !v.empty()) memcpy(out, v.data(), v.size() * sizeof(uint32_t));
    return MPRG_OK;
}

extern "C" int mprg_prg_to_gfa(const char *prg, int64_t length, char *out, int64_t capacity, int64_t *n) {
    if ((!prg && length > 0) || length < 0 || !n) return MPRG_E_BAD_ARG;
    RawBuf<char> s;
    const int rc = prg_to_gfa(prg, length, s);
    if (rc) return rc;
    *n = (int64_t)s.size();
    if (out && capacity >= (int64_t)s.size()) memcpy(out, s.data(), s.size());
    return MPRG_OK;
}

extern "C" int mprg_writer_open(const char *output_prefix, int32_t what, mprg_writer **out) {
    if (!output_prefix || !out || (what & ~15) || (what & 7) == 0) return MPRG_E_BAD_ARG;
    mprg_writer *w = new mprg_writer();
    w->prefix = output_prefix;
    w->what = what;
    *out = w;
    return MPRG_OK;
}

extern "C" const char *mprg_writer_error(const mprg_writer *w) { return w ? w->err.c_str() : ""; }

static int writer_flush_entry(mprg_writer *w, const Encoded &e) {
    if ((w->what & MPRG_WRITE_BIN) &&
        !w->zbin.add(e.name + ".bin", e.bin.data(), e.bin.size() * sizeof(uint32_t), e.crc_bin)) {
        w->err = "cannot write " + w->prefix + ".prg.bin.zip: " + strerror(errno);
        return MPRG_E_INTERNAL;
    }
    if ((w->what & MPRG_WRITE_GFA) && !w->zgfa.add(e.name + ".gfa", e.gfa.data(), e.gfa.size(), e.crc_gfa)) {
        w->err = "cannot write " + w->prefix + ".prg.gfa.zip: " + strerror(errno);
        return MPRG_E_INTERNAL;
    }
    return MPRG_OK;
}

extern "C" int mprg_writer_add(mprg_writer *w, const mprg_result *res, const int32_t *h_loci,
                               const char *const *names, int32_t n, int32_t n_threads) {
    if (!w || !res || n < 0 || (n > 0 && (!h_loci || !names))) return MPRG_E_BAD_ARG;
    std::vector<Encoded> enc((size_t)n);
    const int what = w->what;
    auto encode_one = [&](int i) {
        Encoded &e = enc[(size_t)i];
        e.name = names[i];
        int64_t len = 0;
        const char *prg = mprg_result_prg(res, h_loci[i], &len);
        if (!prg) {
            e.rc = MPRG_E_BAD_ARG;
            return;
        }
        if (what & MPRG_WRITE_BIN) {
            e.rc = encode_prg(prg, len, e.bin);
            if (e.rc) return;
            e.crc_bin = (uint32_t)crc32(0L, (const Bytef *)e.bin.data(), (uInt)(e.bin.size() * sizeof(uint32_t)));
        }
        if (what & MPRG_WRITE_GFA) {
            e.rc = prg_to_gfa(prg, len, e.gfa);
            if (e.rc) return;
            e.crc_gfa = (uint32_t)crc32(0L, (const Bytef *)e.gfa.data(), (uInt)e.gfa.size());
        }
    };
    if (n_threads > 1 && n >= 64 && (w->n_added > 0 || (w->what & MPRG_WRITE_PART) || n > 1)) {
        // Streaming: the archives are appended in locus order by one thread each WHILE the host threads encode
        // (the indices are handed out in increasing order, so entry i is ready about when the appenders get to
        // it); the serial appends -- a third of the writers' time on 8-16 cores -- hide behind the encoding.
        if (!w->zips_open) {
            if ((what & MPRG_WRITE_BIN) && !w->zbin.open_path(w->prefix + ".prg.bin.zip")) {
                w->err = "cannot create " + w->prefix + ".prg.bin.zip: " + strerror(errno);
                return MPRG_E_INTERNAL;
            }
            if ((what & MPRG_WRITE_GFA) && !w->zgfa.open_path(w->prefix + ".prg.gfa.zip")) {
                w->err = "cannot create " + w->prefix + ".prg.gfa.zip: " + strerror(errno);
                return MPRG_E_INTERNAL;
            }
            w->zips_open = true;
            if (w->n_added == 1) {
                const int rc = writer_flush_entry(w, w->first);
                if (rc) return rc;
                w->first = Encoded();
            }
        }
        std::unique_ptr<std::atomic<int>[]> state(new std::atomic<int>[(size_t)n]);  // 0 pending, 1 encoded, 2 failed
        for (int i = 0; i < n; ++i) state[(size_t)i].store(0, std::memory_order_relaxed);
        std::atomic<bool> stop{false};
        int rc_bin = MPRG_OK, rc_gfa = MPRG_OK;
        std::string err_bin, err_gfa;
        auto appender = [&](bool bin) {
            for (int i = 0; i < n; ++i) {
                int st;
                while ((st = state[(size_t)i].load(std::memory_order_acquire)) == 0) {
                    if (stop.load(std::memory_order_relaxed)) return;
                    sched_yield();
                }
                if (st != 1 || stop.load(std::memory_order_relaxed)) return;
                const Encoded &e = enc[(size_t)i];
                const bool ok = bin ? w->zbin.add(e.name + ".bin", e.bin.data(), e.bin.size() * sizeof(uint32_t), e.crc_bin)
                                    : w->zgfa.add(e.name + ".gfa", e.gfa.data(), e.gfa.size(), e.crc_gfa);
                if (!ok) {
                    (bin ? err_bin : err_gfa) = "cannot write " + w->prefix + (bin ? ".prg.bin.zip: " : ".prg.gfa.zip: ") +
                                               strerror(errno);
                    (bin ? rc_bin : rc_gfa) = MPRG_E_INTERNAL;
                    stop.store(true);
                    return;
                }
            }
        };
        std::thread t_bin, t_gfa;
        if (what & MPRG_WRITE_BIN) t_bin = std::thread(appender, true);
        if (what & MPRG_WRITE_GFA) t_gfa = std::thread(appender, false);
        parallel_for(n, n_threads, [&](int i) {
            if (!stop.load(std::memory_order_relaxed)) encode_one(i);
            else enc[(size_t)i].rc = MPRG_E_INTERNAL;  // (another locus failed: nothing more is written)
            const bool bad = enc[(size_t)i].rc != 0;
            if (bad) stop.store(true);
            state[(size_t)i].store(bad ? 2 : 1, std::memory_order_release);
        });
        if (what & MPRG_WRITE_PRG) {
            w->fa.reserve(w->fa.size() + (size_t)n);
            for (int i = 0; i < n && !stop.load(); ++i) {
                int64_t len = 0;
                const char *prg = mprg_result_prg(res, h_loci[i], &len);
                w->fa.emplace_back(enc[(size_t)i].name, std::string(prg, (size_t)len));
            }
        }
        if (t_bin.joinable()) t_bin.join();
        if (t_gfa.joinable()) t_gfa.join();
        if (rc_bin != MPRG_OK || rc_gfa != MPRG_OK) {
            w->err = rc_bin != MPRG_OK ? err_bin : err_gfa;
            return MPRG_E_INTERNAL;
        }
        for (int i = 0; i < n; ++i) {
            const int rc = enc[(size_t)i].rc;
            if (rc && rc != MPRG_E_INTERNAL) {  // the locus that could not be encoded (not the ones skipped after it)
                w->err = "PRG of " + enc[(size_t)i].name + " cannot be encoded";
                return rc;
            }
        }
        if (stop.load()) {
            w->err = "writer stopped";
            return MPRG_E_INTERNAL;
        }
        w->n_added += n;
        return MPRG_OK;
    }
    parallel_for(n, n_threads, encode_one);
    for (int i = 0; i < n; ++i)
        if (enc[(size_t)i].rc) {
            w->err = "PRG of " + enc[(size_t)i].name + " cannot be encoded";
            return enc[(size_t)i].rc;
        }
    if (n == 0) return MPRG_OK;
    // the only locus of a run is kept back: one locus => plain files, no archives
    int begin = 0;
    if (w->n_added == 0 && n == 1 && !(w->what & MPRG_WRITE_PART)) {
        if (what & MPRG_WRITE_PRG) {
            int64_t len = 0;
            const char *prg = mprg_result_prg(res, h_loci[0], &len);
            w->fa.emplace_back(enc[0].name, std::string(prg, (size_t)len));
        }
        w->first = std::move(enc[0]);
        w->n_added = 1;
        return MPRG_OK;
    }
    if (!w->zips_open) {
        if ((what & MPRG_WRITE_BIN) && !w->zbin.open_path(w->prefix + ".prg.bin.zip")) {
            w->err = "cannot create " + w->prefix + ".prg.bin.zip: " + strerror(errno);
            return MPRG_E_INTERNAL;
        }
        if ((what & MPRG_WRITE_GFA) && !w->zgfa.open_path(w->prefix + ".prg.gfa.zip")) {
            w->err = "cannot create " + w->prefix + ".prg.gfa.zip: " + strerror(errno);
            return MPRG_E_INTERNAL;
        }
        w->zips_open = true;
        if (w->n_added == 1) {
            const int rc = writer_flush_entry(w, w->first);
            if (rc) return rc;
            w->first = Encoded();
        }
    }
    // the two archives and the .prg.fa records are independent streams: one host thread each
    int rc_bin = MPRG_OK, rc_gfa = MPRG_OK;
    std::string err_bin, err_gfa;
    auto append_bin = [&]() {
        for (int i = begin; i < n && rc_bin == MPRG_OK; ++i) {
            const Encoded &e = enc[(size_t)i];
            if (!w->zbin.add(e.name + ".bin", e.bin.data(), e.bin.size() * sizeof(uint32_t), e.crc_bin)) {
                err_bin = "cannot write " + w->prefix + ".prg.bin.zip: " + strerror(errno);
                rc_bin = MPRG_E_INTERNAL;
            }
        }
    };
    auto append_gfa = [&]() {
        for (int i = begin; i < n && rc_gfa == MPRG_OK; ++i) {
            const Encoded &e = enc[(size_t)i];
            if (!w->zgfa.add(e.name + ".gfa", e.gfa.data(), e.gfa.size(), e.crc_gfa)) {
                err_gfa = "cannot write " + w->prefix + ".prg.gfa.zip: " + strerror(errno);
                rc_gfa = MPRG_E_INTERNAL;
            }
        }
    };
    std::thread t_bin, t_gfa;
    const bool threaded = n_threads > 1 && n >= 64;
    if (what & MPRG_WRITE_BIN) {
        if (threaded) t_bin = std::thread(append_bin);
        else append_bin();
    }
    if (what & MPRG_WRITE_GFA) {
        if (threaded) t_gfa = std::thread(append_gfa);
        else append_gfa();
    }
    if (what & MPRG_WRITE_PRG) {
        w->fa.reserve(w->fa.size() + (size_t)n);
        for (int i = begin; i < n; ++i) {
            int64_t len = 0;
            const char *prg = mprg_result_prg(res, h_loci[i], &len);
            w->fa.emplace_back(enc[(size_t)i].name, std::string(prg, (size_t)len));
        }
    }
    if (t_bin.joinable()) t_bin.join();
    if (t_gfa.joinable()) t_gfa.join();
    if (rc_bin != MPRG_OK || rc_gfa != MPRG_OK) {
        w->err = rc_bin != MPRG_OK ? err_bin : err_gfa;
        return MPRG_E_INTERNAL;
    }
    w->n_added += n;
    return MPRG_OK;
}

// ---- update_DS archive (prg_builder.py:145-147, input_output_files.py:95-104) ---------------------------------
// One member per locus, named by the locus, holding what a PrgBuilder is made of as TABLES instead of a pickle
// of Python objects: the pre-order node table, the row subsets, the record titles, the root alignment (4-bit
// packed rows) and the PRG string.  make_prg_b200.prg_builder.PrgBuilder.deserialize_from_bytes builds the
// objects on load; make_prg_b200/utils/reference_export.py turns the archive into the reference's pickles
// where Biopython is installed.  Layout (little endian):
//   char magic[8] = "MPRGDS01"
//   int32 max_nesting, min_match_length, n_rows, n_cols, n_nodes, n_sites, packed_stride, reserved
//   int64 pool_len, titles_len, prg_len
//   int32 kind[n], parent[n], nesting_level[n], c0[n], c1[n], n_rows[n], n_children[n]; int64 row_off[n]
//   int32 pool[pool_len]; titles (joined by '\n'); packed rows (n_rows * packed_stride); prg
extern "C" int mprg_writer_add_ds(mprg_writer *w, const mprg_result *res, const mprg_msa_set *msas,
                                  const int32_t *h_loci, const char *const *names, int32_t n, int32_t max_nesting,
                                  int32_t min_match_length, int32_t n_threads) {
    if (!w || !res || !msas || n < 0 || (n > 0 && (!h_loci || !names))) return MPRG_E_BAD_ARG;
    if (n == 0) return MPRG_OK;
    const bool avx2 = __builtin_cpu_supports("avx2") && !getenv("MPRG_NO_AVX2");
    // Every member is header + tables (small, built per locus) followed by the packed rows and the PRG, which
    // are written straight from where they are: sizes first, then every locus is checksummed and written at its
    // own offset by the host threads in parallel (pwrite), the archive's directory entries in locus order.
    struct Member {
        std::string head;  // header + node table + row pool + titles
        const uint8_t *rows = nullptr;
        std::vector<uint8_t> packed_here;  // rows packed now (text loader)
        size_t rows_bytes = 0;
        const char *prg = nullptr;
        size_t prg_bytes = 0;
        uint32_t crc = 0;
        uint64_t offset = 0;
        int rc = MPRG_OK;
        size_t size() const { return head.size() + rows_bytes + prg_bytes; }
    };
    std::vector<Member> mem((size_t)n);
    parallel_for(n, n_threads, [&](int i) {
        Member &m = mem[(size_t)i];
        const int l = h_loci[i];
        if (l < 0 || l >= msas->n || l >= mprg_result_n_loci(res)) {
            m.rc = MPRG_E_BAD_ARG;
            return;
        }
        const int32_t n_nodes = mprg_result_n_nodes(res, l);
        const int64_t pool_len = mprg_result_row_pool_size(res, l);
        const int32_t rows = msas->n_rows[(size_t)l], cols = msas->n_cols[(size_t)l];
        const int64_t stride = packed_stride(cols);
        int64_t prg_len = 0;
        m.prg = mprg_result_prg(res, l, &prg_len);
        m.prg_bytes = (size_t)prg_len;
        const std::string &titles = msas->titles[(size_t)l];
        const size_t header = 8 + 8 * 4 + 3 * 8;
        m.head.resize(header + (size_t)n_nodes * (7 * 4 + 8) + (size_t)pool_len * 4 + titles.size());
        char *p = &m.head[0];
        memcpy(p, "MPRGDS01", 8);
        int32_t h32[8] = {max_nesting, min_match_length, rows, cols, n_nodes, mprg_result_n_sites(res, l), (int32_t)stride, 0};
        memcpy(p + 8, h32, sizeof(h32));
        int64_t h64[3] = {pool_len, (int64_t)titles.size(), prg_len};
        memcpy(p + 8 + sizeof(h32), h64, sizeof(h64));
        char *q = p + header;
        std::vector<int32_t> cols32((size_t)std::max(n_nodes, 1) * 7);
        std::vector<int64_t> roff((size_t)std::max(n_nodes, 1));
        int32_t *kind = cols32.data(), *parent = kind + n_nodes, *level = parent + n_nodes, *c0 = level + n_nodes;
        int32_t *c1 = c0 + n_nodes, *nr = c1 + n_nodes, *nch = nr + n_nodes;
        if (n_nodes > 0 && mprg_result_nodes(res, l, kind, parent, level, c0, c1, nr, roff.data(), nch) != MPRG_OK) {
            m.rc = MPRG_E_INTERNAL;
            return;
        }
        memcpy(q, cols32.data(), (size_t)n_nodes * 7 * 4);
        q += (size_t)n_nodes * 7 * 4;
        memcpy(q, roff.data(), (size_t)n_nodes * 8);
        q += (size_t)n_nodes * 8;
        if (pool_len > 0) {
            std::vector<int32_t> pool((size_t)pool_len);
            mprg_result_row_pool(res, l, pool.data());
            memcpy(q, pool.data(), (size_t)pool_len * 4);
            q += (size_t)pool_len * 4;
        }
        memcpy(q, titles.data(), titles.size());
        m.rows_bytes = (size_t)rows * (size_t)stride;
        if (m.rows_bytes) {
            if (msas->packed) {
                m.rows = msas->packed + msas->packed_offsets[(size_t)l];
            } else {
                m.packed_here.resize(m.rows_bytes);
                pack_matrix(msas->ascii + msas->offsets[(size_t)l], rows, cols, m.packed_here.data(), avx2);
                m.rows = m.packed_here.data();
            }
        }
        uLong crc = crc32(0L, (const Bytef *)m.head.data(), (uInt)m.head.size());
        for (size_t at = 0; at < m.rows_bytes; at += (size_t)1 << 30)
            crc = crc32(crc, (const Bytef *)m.rows + at, (uInt)std::min<size_t>(m.rows_bytes - at, (size_t)1 << 30));
        for (size_t at = 0; at < m.prg_bytes; at += (size_t)1 << 30)
            crc = crc32(crc, (const Bytef *)m.prg + at, (uInt)std::min<size_t>(m.prg_bytes - at, (size_t)1 << 30));
        m.crc = (uint32_t)crc;
    });
    for (int i = 0; i < n; ++i) {
        if (mem[(size_t)i].rc != MPRG_OK) {
            w->err = std::string("cannot serialise the update data of ") + names[i];
            return mem[(size_t)i].rc;
        }
        if (mem[(size_t)i].size() > 0xFFFFFFFEull) {
            w->err = std::string("update data of ") + names[i] + " exceeds 4 GiB";
            return MPRG_E_INTERNAL;
        }
    }
    if (!w->ds_open) {
        if (!w->zds.open_path(w->prefix + ".update_DS.zip")) {
            w->err = "cannot create " + w->prefix + ".update_DS.zip: " + strerror(errno);
            return MPRG_E_INTERNAL;
        }
        w->ds_open = true;
    }
    // local headers in order (small), data regions left as holes and filled in parallel
    ZipFile &z = w->zds;
    if (fflush(z.f) != 0) {
        w->err = "cannot write " + w->prefix + ".update_DS.zip: " + strerror(errno);
        return MPRG_E_INTERNAL;
    }
    const int fd = fileno(z.f);
    std::vector<std::string> heads((size_t)n);
    uint64_t pos = z.pos;
    for (int i = 0; i < n; ++i) {
        Member &m = mem[(size_t)i];
        const std::string name = names[i];
        if (name.size() > 0xFFFF) {
            w->err = "locus name too long: " + name;
            return MPRG_E_INTERNAL;
        }
        std::string &h = heads[(size_t)i];
        ZipFile::le32(h, 0x04034b50u);
        ZipFile::le16(h, 20);
        ZipFile::le16(h, 0x0800);
        ZipFile::le16(h, 0);
        ZipFile::le16(h, z.dos_time);
        ZipFile::le16(h, z.dos_date);
        ZipFile::le32(h, m.crc);
        ZipFile::le32(h, (uint32_t)m.size());
        ZipFile::le32(h, (uint32_t)m.size());
        ZipFile::le16(h, (uint16_t)name.size());
        ZipFile::le16(h, 0);
        h += name;
        z.entries.push_back({name, m.crc, (uint32_t)m.size(), pos});
        m.offset = pos;
        pos += h.size() + m.size();
    }
    std::atomic<int> failed{0};
    auto put = [&](const void *data, size_t len, uint64_t at) {
        const char *p = (const char *)data;
        while (len > 0) {
            const ssize_t k = pwrite(fd, p, len, (off_t)at);
            if (k <= 0) {
                if (k < 0 && errno == EINTR) continue;
                failed = errno ? errno : EIO;
                return;
            }
            p += k;
            at += (uint64_t)k;
            len -= (size_t)k;
        }
    };
    parallel_for(n, n_threads, [&](int i) {
        const Member &m = mem[(size_t)i];
        uint64_t at = m.offset;
        put(heads[(size_t)i].data(), heads[(size_t)i].size(), at);
        at += heads[(size_t)i].size();
        put(m.head.data(), m.head.size(), at);
        at += m.head.size();
        if (m.rows_bytes) put(m.rows, m.rows_bytes, at);
        at += m.rows_bytes;
        if (m.prg_bytes) put(m.prg, m.prg_bytes, at);
    });
    if (failed.load() || fseeko(z.f, (off_t)pos, SEEK_SET) != 0) {
        w->err = "cannot write " + w->prefix + ".update_DS.zip: " + strerror(failed.load() ? failed.load() : errno);
        return MPRG_E_INTERNAL;
    }
    z.pos = pos;
    w->n_ds += n;
    return MPRG_OK;
}

extern "C" int mprg_writer_close(mprg_writer *w, int64_t *n_loci, int64_t *bytes_written) {
    if (!w) return MPRG_E_BAD_ARG;
    int rc = MPRG_OK;
    int64_t bytes = 0;
    if (w->n_added > 0) {
        if (w->what & MPRG_WRITE_PRG) {
            // input_output_files.py:86-92: the per-locus <name>.prg.fa files are concatenated in sorted order
            std::vector<size_t> order(w->fa.size());
            for (size_t i = 0; i < order.size(); ++i) order[i] = i;
            std::vector<std::string> keys(w->fa.size());
            for (size_t i = 0; i < keys.size(); ++i) keys[i] = w->fa[i].first + ".prg.fa";
            std::stable_sort(order.begin(), order.end(), [&](size_t a, size_t b) { return keys[a] < keys[b]; });
            std::string text;
            for (size_t k : order) {
                text.push_back('>');
                text += w->fa[k].first;
                text.push_back('\n');
                text += w->fa[k].second;
                text.push_back('\n');
            }
            if (!write_whole(w->prefix + ".prg.fa", text.data(), text.size())) {
                w->err = "cannot write " + w->prefix + ".prg.fa: " + strerror(errno);
                rc = MPRG_E_INTERNAL;
            }
            bytes += (int64_t)text.size();
        }
        if (w->n_added == 1 && !w->zips_open) {
            const Encoded &e = w->first;
            if ((w->what & MPRG_WRITE_BIN) &&
                !write_whole(w->prefix + ".prg.bin", e.bin.data(), e.bin.size() * sizeof(uint32_t))) {
                w->err = "cannot write " + w->prefix + ".prg.bin: " + strerror(errno);
                rc = MPRG_E_INTERNAL;
            }
            if ((w->what & MPRG_WRITE_GFA) && !write_whole(w->prefix + ".prg.gfa", e.gfa.data(), e.gfa.size())) {
                w->err = "cannot write " + w->prefix + ".prg.gfa: " + strerror(errno);
                rc = MPRG_E_INTERNAL;
            }
            bytes += (int64_t)(e.bin.size() * sizeof(uint32_t) + e.gfa.size());
        } else {
            if (w->what & MPRG_WRITE_BIN) {
                if (!w->zbin.finish()) rc = MPRG_E_INTERNAL;
                bytes += (int64_t)w->zbin.pos;
            }
            if (w->what & MPRG_WRITE_GFA) {
                if (!w->zgfa.finish()) rc = MPRG_E_INTERNAL;
                bytes += (int64_t)w->zgfa.pos;
            }
            if (rc && w->err.empty()) w->err = "cannot finish the archives of " + w->prefix;
        }
    }
    if (w->ds_open) {
        if (!w->zds.finish()) {
            rc = MPRG_E_INTERNAL;
            if (w->err.empty()) w->err = "cannot finish " + w->prefix + ".update_DS.zip";
        }
        bytes += (int64_t)w->zds.pos;
        w->ds_open = false;
    }
    if (n_loci) *n_loci = w->n_added;
    if (bytes_written) *bytes_written = bytes;
    if (rc == MPRG_OK) delete w;  // on failure the caller reads mprg_writer_error, then mprg_writer_abort
    return rc;
}

extern "C" void mprg_writer_abort(mprg_writer *w) {
    if (!w) return;
    w->zbin.discard();  // closes and removes the unfinished archives
    w->zgfa.discard();
    w->zds.discard();
    delete w;
}

// Final files of a run that was built as several parts (one per GPU shard, each written by its own
// writer opened with MPRG_WRITE_PART): .prg.fa records merged in sorted order (input_output_files.py:86-92),
// archive members concatenated part by part; a run of a single locus gets plain .prg.bin / .prg.gfa files
// (input_output_files.py:113-135).  Parts that do not exist hold no locus.  The parts are removed afterwards.
extern "C" int mprg_merge_outputs(const char *const *part_prefixes, int32_t n_parts, const char *output_prefix,
                                  int32_t what, int64_t *n_loci, char *err_buf, int64_t err_capacity) {
    if (!part_prefixes || n_parts < 0 || !output_prefix || (what & ~(7 | MPRG_WRITE_DS)) || what == 0) return MPRG_E_BAD_ARG;
    std::string err;
    auto fail = [&](const std::string &m) {
        if (err_buf && err_capacity > 0) {
            const size_t k = std::min<size_t>(m.size(), (size_t)err_capacity - 1);
            memcpy(err_buf, m.data(), k);
            err_buf[k] = 0;
        }
        return MPRG_E_INTERNAL;
    };
    auto exists = [](const std::string &p) { return access(p.c_str(), R_OK) == 0; };
    const std::string out = output_prefix;
    int64_t total = -1;
    // ---- archives ----
    struct Kind { int bit; const char *zip_ext, *plain_ext; };
    const Kind kinds[2] = {{MPRG_WRITE_BIN, ".prg.bin.zip", ".prg.bin"}, {MPRG_WRITE_GFA, ".prg.gfa.zip", ".prg.gfa"}};
    for (const Kind &k : kinds) {
        if (!(what & k.bit)) continue;
        std::vector<std::string> parts;
        for (int i = 0; i < n_parts; ++i) {
            const std::string p = std::string(part_prefixes[i]) + k.zip_ext;
            if (exists(p)) parts.push_back(p);
        }
        // one member in all => plain file
        int64_t members = 0;
        std::string only_z;
        std::vector<ZipEntry> only_e;
        for (const std::string &p : parts) {
            std::string z;
            std::vector<ZipEntry> e;
            uint64_t cd = 0;
            if (!read_whole(p, z) || !parse_stored_zip(z, e, cd)) return fail("cannot read the archive part " + p);
            members += (int64_t)e.size();
            if (members == (int64_t)e.size() && e.size() == 1) {
                only_z.swap(z);
                only_e = e;
            }
            if (members > 1) break;
        }
        if (members == 1) {
            const char *d = nullptr;
            size_t n = 0;
            if (!member_data(only_z, only_e[0], d, n) || !write_whole(out + k.plain_ext, d, n))
                return fail("cannot write " + out + k.plain_ext);
            total = 1;
        } else if (members > 1) {
            ZipFile zf;
            if (!zf.open_path(out + k.zip_ext)) return fail("cannot create " + out + k.zip_ext + ": " + strerror(errno));
            for (const std::string &p : parts)
                if (!zf.append_archive(p, err)) {
                    zf.discard();
                    return fail(err);
                }
            total = (int64_t)zf.entries.size();
            if (!zf.finish()) return fail("cannot finish " + out + k.zip_ext);
        }
        for (const std::string &p : parts) unlink(p.c_str());
    }
    // ---- update_DS archives: always an archive, members appended part by part ----
    if (what & MPRG_WRITE_DS) {
        std::vector<std::string> parts;
        for (int i = 0; i < n_parts; ++i) {
            const std::string p = std::string(part_prefixes[i]) + ".update_DS.zip";
            if (exists(p)) parts.push_back(p);
        }
        if (!parts.empty()) {
            ZipFile zf;
            if (!zf.open_path(out + ".update_DS.zip")) return fail("cannot create " + out + ".update_DS.zip: " + strerror(errno));
            for (const std::string &p : parts)
                if (!zf.append_archive(p, err)) {
                    zf.discard();
                    return fail(err);
                }
            if (!zf.finish()) return fail("cannot finish " + out + ".update_DS.zip");
            for (const std::string &p : parts) unlink(p.c_str());
        }
    }
    // ---- .prg.fa: every part is sorted by "<name>.prg.fa"; merge ----
    if (what & MPRG_WRITE_PRG) {
        struct Rec { std::string key; const char *p; size_t n; };
        std::vector<std::string> texts((size_t)n_parts);
        std::vector<Rec> recs;
        for (int i = 0; i < n_parts; ++i) {
            const std::string p = std::string(part_prefixes[i]) + ".prg.fa";
            if (!exists(p)) continue;
            if (!read_whole(p, texts[(size_t)i])) return fail("cannot read " + p);
            const std::string &t = texts[(size_t)i];
            size_t at = 0;
            while (at < t.size()) {  // ">name\nprg\n"
                const size_t e1 = t.find('\n', at);
                if (t[at] != '>' || e1 == std::string::npos) return fail("malformed part " + p);
                size_t e2 = t.find('\n', e1 + 1);
                if (e2 == std::string::npos) e2 = t.size() - 1;
                recs.push_back(Rec{t.substr(at + 1, e1 - at - 1) + ".prg.fa", t.data() + at, e2 + 1 - at});
                at = e2 + 1;
            }
        }
        std::stable_sort(recs.begin(), recs.end(), [](const Rec &a, const Rec &b) { return a.key < b.key; });
        if (!recs.empty()) {
            std::string text;
            size_t bytes = 0;
            for (const Rec &r : recs) bytes += r.n;
            text.reserve(bytes);
            for (const Rec &r : recs) text.append(r.p, r.n);
            if (!write_whole(out + ".prg.fa", text.data(), text.size())) return fail("cannot write " + out + ".prg.fa");
        }
        total = (int64_t)recs.size();
        for (int i = 0; i < n_parts; ++i) unlink((std::string(part_prefixes[i]) + ".prg.fa").c_str());
    }
    if (n_loci) *n_loci = std::max<int64_t>(total, 0);
    return MPRG_OK;
}
