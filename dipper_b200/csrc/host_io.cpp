// Host-side pieces of dipper_b200 (include/dipper_host.h): encoders, Newick writer /
// backbone reader.  Plain C++17; no CUDA here.
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <algorithm>
#include <cctype>
#include <cmath>
#include <cstdint>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <vector>
#include "../../include/dipper_b200.h"
#include "../../include/dipper_host.h"

namespace dipb {
void set_error(const char* fmt, ...);
}

namespace {

struct Code4 {
    unsigned char t[256];
    Code4() {
        for (int i = 0; i < 256; i++) t[i] = 4;
        t['A'] = 0; t['C'] = 1; t['G'] = 2; t['T'] = 3; t['U'] = 3;
    }
};
const Code4 kCode4;
struct Code2 {   // twoBitCompressor: anything but A C G T/U packs as 0
    unsigned char t[256];
    Code2() {
        for (int i = 0; i < 256; i++) t[i] = 0;
        t['C'] = 1; t['G'] = 2; t['T'] = 3; t['U'] = 3;
    }
};
const Code2 kCode2;

// "%g" (6 significant digits; what ostream << double prints) for 1e-15 <= |v| < 1e6 and 0, digit for digit what glibc's
// printf writes: the value is M * 2^-sh exactly, so round-half-even on M * 10^p >> sh in 128-bit integers is the correctly
// rounded 6-digit decimal.  Returns 0 when the caller has to fall back to snprintf (other magnitudes, inf, nan).  4.4x faster
// than snprintf; a 30 000-tip Newick string holds 60 000 branch lengths (tests/test_abi.py compares both on random values).
static inline int fmt_g6_fast(double v, char* out) {
    char* o = out;
    if (v == 0.0) { if (std::signbit(v)) *o++ = '-'; *o++ = '0'; return (int)(o - out); }
    if (!(v == v) || std::isinf(v)) return 0;
    if (v < 0) { *o++ = '-'; v = -v; }
    if (!(v >= 1e-15 && v < 1e6)) return 0;
    uint64_t bits;
    memcpy(&bits, &v, 8);
    const int be = (int)((bits >> 52) & 0x7ff);           // normal numbers only (the range check above excludes denormals)
    const uint64_t M = (bits & ((1ull << 52) - 1)) | (1ull << 52);   // v = M * 2^(be - 1075)
    const int sh = 1075 - be;
    static const uint64_t P10[23] = {1ull, 10ull, 100ull, 1000ull, 10000ull, 100000ull, 1000000ull, 10000000ull, 100000000ull, 1000000000ull,
                                     10000000000ull, 100000000000ull, 1000000000000ull, 10000000000000ull, 100000000000000ull,
                                     1000000000000000ull, 10000000000000000ull, 100000000000000000ull, 1000000000000000000ull,
                                     10000000000000000000ull, 0, 0, 0};
    int X = (int)(((be - 1023) * 78913) >> 18);            // floor(log10(v)) or one less (78913 / 2^18 = log10(2)); fixed up below
    uint64_t D = 0;
    for (int tries = 0; tries < 3; tries++) {
        const int p = 5 - X;                              // digits = round(v * 10^p)
        if (p < 0 || p > 19 || sh <= 0 || sh > 120) return 0;
        const unsigned __int128 N = (unsigned __int128)M * P10[p];
        const unsigned __int128 q = N >> sh, rem = N & (((unsigned __int128)1 << sh) - 1), half = (unsigned __int128)1 << (sh - 1);
        D = (uint64_t)q;
        if (rem > half || (rem == half && (D & 1ull))) D++;
        if (D < 100000ull) { X--; continue; }
        if (D == 1000000ull) { D = 100000ull; X++; break; }
        if (D > 1000000ull) { X++; continue; }
        break;
    }
    if (D < 100000ull || D > 999999ull || X >= 6) return 0;
    char dig[6];
    for (int i = 5; i >= 0; i--) { dig[i] = (char)('0' + D % 10); D /= 10; }
    int nd = 6;
    while (nd > 1 && dig[nd - 1] == '0') nd--;            // %g strips trailing zeros
    if (X < -4) {                                         // scientific
        *o++ = dig[0];
        if (nd > 1) { *o++ = '.'; memcpy(o, dig + 1, (size_t)(nd - 1)); o += nd - 1; }
        *o++ = 'e'; *o++ = '-';
        const int ax = -X;
        *o++ = (char)('0' + ax / 10); *o++ = (char)('0' + ax % 10);
        return (int)(o - out);
    }
    if (X >= 0) {
        const int ip = X + 1;                             // digits before the point
        for (int i = 0; i < ip; i++) *o++ = i < nd ? dig[i] : '0';
        if (nd > ip) { *o++ = '.'; memcpy(o, dig + ip, (size_t)(nd - ip)); o += nd - ip; }
    } else {
        *o++ = '0'; *o++ = '.';
        for (int i = 0; i < -X - 1; i++) *o++ = '0';
        memcpy(o, dig, (size_t)nd); o += nd;
    }
    return (int)(o - out);
}

void append_g(std::string& s, double v) {
    char buf[64];
    int n = fmt_g6_fast(v, buf);
    if (n == 0) n = snprintf(buf, sizeof buf, "%g", v);  // ostream<<double default formatting
    s.append(buf, (size_t)n);
}

char* dup_string(const std::string& s) {
    char* p = (char*)malloc(s.size() + 1);
    if (!p) return nullptr;
    memcpy(p, s.data(), s.size());
    p[s.size()] = 0;
    return p;
}

}  // namespace

extern "C" {

void dipb_pack4(const char* seq, size_t len, uint64_t* out) {
    size_t nw = (len + 15) / 16;
    for (size_t w = 0; w < nw; w++) {
        uint64_t v = 0;
        size_t end = w * 16 + 16 < len ? w * 16 + 16 : len;
        for (size_t s = w * 16; s < end; s++) v |= (uint64_t)kCode4.t[(unsigned char)seq[s]] << (4 * (s & 15));
        out[w] = v;
    }
}

void dipb_pack2(const char* seq, size_t len, uint64_t* out) {
    size_t nw = (len + 31) / 32;
    for (size_t w = 0; w < nw; w++) {
        uint64_t v = 0;
        size_t end = w * 32 + 32 < len ? w * 32 + 32 : len;
        for (size_t s = w * 32; s < end; s++) {
            uint64_t c = kCode4.t[(unsigned char)seq[s]];
            v |= (c & 3 & (uint64_t)-(int64_t)(c < 4)) << (2 * (s & 31));
        }
        out[w] = v;
    }
}

// ---- FASTA ingest (dipper_host.h) -----------------------------------------------------------------
}  // extern "C"

struct dipb_fasta {
    std::vector<std::string> names;
    std::vector<uint64_t> lens, word_off;
    uint64_t* words = nullptr;      // malloc'd, not zero-filled: the packing threads write (and first-touch) every word
    ~dipb_fasta() { free(words); }
};

namespace {
// isgraph() in the C locale, branch-free so that the counting loop vectorises
inline bool is_graphic(unsigned char c) { return (unsigned)(c - 33) < 94u; }
inline uint64_t count_graphic(const unsigned char* p, const unsigned char* end) {
    uint64_t n = 0;
    while (end - p >= 32) {                     // byte lanes: vectorises (psubb / pcmpgtb / psadbw); FASTA lines are 60-100 bytes
        unsigned char acc = 0;
        for (int i = 0; i < 32; i++) acc += (unsigned char)((unsigned char)(p[i] - 33) < 94);
        n += acc;
        p += 32;
    }
    for (; p < end; p++) n += (unsigned)(*p - 33) < 94u;
    return n;
}
template <class F>
void run_threads(unsigned nt, size_t n, F fn) {   // contiguous blocks of [0, n)
    if (nt <= 1 || n < 2) { fn((size_t)0, n); return; }
    std::vector<std::thread> th;
    const size_t per = (n + nt - 1) / nt;
    for (unsigned t = 0; t < nt; t++) {
        const size_t a = std::min(n, t * per), b = std::min(n, a + per);
        if (a < b) th.emplace_back([=]() { fn(a, b); });
    }
    for (auto& x : th) x.join();
}
}  // namespace

extern "C" {

int dipb_fasta_open(const char* path, int bits, int threads, dipb_fasta** out) {
    if (!path || !out || (bits != 2 && bits != 4)) { dipb::set_error("dipb_fasta_open: bad argument"); return DIPB_E_ARG; }
    const int fd = open(path, O_RDONLY);
    if (fd < 0) { dipb::set_error("dipb_fasta_open: cannot open %s", path); return DIPB_E_ARG; }
    struct stat sb;
    if (fstat(fd, &sb) != 0) { close(fd); dipb::set_error("dipb_fasta_open: cannot stat %s", path); return DIPB_E_ARG; }
    const size_t size = (size_t)sb.st_size;
    dipb_fasta* f = new dipb_fasta();
    if (size == 0) { close(fd); f->word_off.push_back(0); *out = f; return 0; }
    const unsigned char* d = (const unsigned char*)mmap(nullptr, size, PROT_READ, MAP_PRIVATE, fd, 0);
    close(fd);
    if (d == MAP_FAILED) { delete f; dipb::set_error("dipb_fasta_open: mmap of %s failed", path); return DIPB_E_NOMEM; }
    if (size >= 2 && d[0] == 0x1f && d[1] == 0x8b) {
        munmap((void*)d, size); delete f;
        dipb::set_error("dipb_fasta_open: %s is gzip-compressed", path);
        return DIPB_E_UNSUPPORTED;
    }
    madvise((void*)d, size, MADV_SEQUENTIAL);
    const bool prof = getenv("DIPB_FASTA_PROFILE") != nullptr;
    auto tp = std::chrono::steady_clock::now();
    auto lap = [&](const char* what) {
        auto now = std::chrono::steady_clock::now();
        if (prof) fprintf(stderr, "[fasta] %s %.1f ms\n", what, std::chrono::duration<double, std::milli>(now - tp).count());
        tp = now;
    };
    unsigned nt = threads > 0 ? (unsigned)threads : std::max(1u, std::thread::hardware_concurrency());
    if (size < (1u << 20)) nt = 1;
    // pass 0: record starts ('>' in column 0), found chunk by chunk
    std::vector<std::vector<size_t>> part(nt);
    {
        std::vector<std::thread> th;
        const size_t per = (size + nt - 1) / nt;
        for (unsigned t = 0; t < nt; t++)
            th.emplace_back([&, t]() {
                const size_t a = std::min(size, t * per), b = std::min(size, a + per);
                const unsigned char* p = d + a;
                while (p < d + b) {
                    const unsigned char* q = (const unsigned char*)memchr(p, '>', (size_t)(d + b - p));
                    if (!q) break;
                    if (q == d || q[-1] == '\n') part[t].push_back((size_t)(q - d));
                    p = q + 1;
                }
            });
        for (auto& x : th) x.join();
    }
    lap("record search");
    std::vector<size_t> start;
    for (auto& v : part) start.insert(start.end(), v.begin(), v.end());
    const size_t n = start.size();
    start.push_back(size);
    f->names.resize(n); f->lens.assign(n, 0); f->word_off.assign(n + 1, 0);
    std::vector<size_t> body(n);
    // pass 1: names and lengths
    run_threads(nt, n, [&](size_t a, size_t b) {
        for (size_t i = a; i < b; i++) {
            const unsigned char* h = d + start[i] + 1;
            const unsigned char* end = d + start[i + 1];
            const unsigned char* nl = (const unsigned char*)memchr(h, '\n', (size_t)(end - h));
            const unsigned char* hend = nl ? nl : end;
            const unsigned char* w = h;
            while (w < hend && !isspace(*w)) w++;
            f->names[i].assign((const char*)h, (size_t)(w - h));
            const unsigned char* s0 = nl ? nl + 1 : end;
            body[i] = (size_t)(s0 - d);
            const uint64_t len = count_graphic(s0, end);
            f->lens[i] = len;
        }
    });
    lap("names + lengths");
    const unsigned per_word = bits == 4 ? 16 : 32;
    for (size_t i = 0; i < n; i++) f->word_off[i + 1] = f->word_off[i] + (f->lens[i] + per_word - 1) / per_word;
    f->words = (uint64_t*)malloc(sizeof(uint64_t) * (size_t)(f->word_off[n] + 1));
    if (!f->words) { munmap((void*)d, size); delete f; dipb::set_error("dipb_fasta_open: out of memory"); return DIPB_E_NOMEM; }
    lap("output allocation");
    // pass 2: pack (same code tables as dipb_pack4 / dipb_pack2)
    run_threads(nt, n, [&](size_t a, size_t b) {
        for (size_t i = a; i < b; i++) {
            uint64_t* o = f->words + f->word_off[i];
            const unsigned char* end = d + start[i + 1];
            uint64_t v = 0, k = 0;
            // line by line: a line that is all graphic characters (the usual case) is packed without per-byte tests
            const unsigned char* p = d + body[i];
            while (p < end) {
                const unsigned char* nl = (const unsigned char*)memchr(p, '\n', (size_t)(end - p));
                const unsigned char* q = nl ? nl : end;
                const bool clean = count_graphic(p, q) == (uint64_t)(q - p);
                const unsigned char* r = p;
                if (clean) {
                    // whole words' worth of characters at a time: independent table look-ups, then one shift into place
                    const unsigned sh = bits == 4 ? 4u : 2u;
                    for (; r + per_word <= q; r += per_word) {
                        uint64_t w = 0;
                        if (bits == 4) {
#pragma GCC unroll 16
                            for (unsigned j = 0; j < 16; j++) w |= (uint64_t)kCode4.t[r[j]] << (4 * j);
                        } else {
#pragma GCC unroll 32
                            for (unsigned j = 0; j < 32; j++) w |= (uint64_t)kCode2.t[r[j]] << (2 * j);
                        }
                        const unsigned ph = (unsigned)(k & (per_word - 1)) * sh;   // bit position inside the open word
                        v |= w << ph;
                        *o++ = v;
                        v = ph ? w >> (64 - ph) : 0;
                        k += per_word;
                    }
                }
                for (; r < q; r++) {
                    if (!clean && !is_graphic(*r)) continue;
                    const uint64_t c = kCode4.t[*r];
                    if (bits == 4) v |= c << (4 * (k & 15));
                    else v |= (c & 3 & (uint64_t)-(int64_t)(c < 4)) << (2 * (k & 31));
                    k++;
                    if ((k & (per_word - 1)) == 0) { *o++ = v; v = 0; }
                }
                p = q + 1;
            }
            if (k & (per_word - 1)) *o = v;
        }
    });
    lap("pack");
    munmap((void*)d, size);
    lap("munmap");
    *out = f;
    return 0;
}
size_t dipb_fasta_count(const dipb_fasta* f) { return f ? f->names.size() : 0; }
const char* dipb_fasta_name(const dipb_fasta* f, size_t i) { return (f && i < f->names.size()) ? f->names[i].c_str() : ""; }
const uint64_t* dipb_fasta_lengths(const dipb_fasta* f) { return f ? f->lens.data() : nullptr; }
const uint64_t* dipb_fasta_word_offsets(const dipb_fasta* f) { return f ? f->word_off.data() : nullptr; }
const uint64_t* dipb_fasta_words(const dipb_fasta* f) { return f ? f->words : nullptr; }
void dipb_fasta_close(dipb_fasta* f) { delete f; }

char* dipb_nj_newick(int n, const int32_t* c0, const int32_t* c1, const double* l0, const double* l1,
                     const char* const* names) {
    std::string s;
    s.reserve((size_t)n * 24);
    struct Frame { int node, stage; };
    std::vector<Frame> st;
    st.push_back({2 * n - 2, 0});
    while (!st.empty()) {
        Frame& f = st.back();
        if (f.node < n) { s += names[f.node]; st.pop_back(); continue; }
        int k = f.node - n;
        if (f.stage == 0) { f.stage = 1; s += '('; st.push_back({c0[k], 0}); }
        else if (f.stage == 1) { f.stage = 2; s += ':'; append_g(s, l0[k]); s += ','; st.push_back({c1[k], 0}); }
        else { s += ':'; append_g(s, l1[k]); s += ')'; st.pop_back(); }
    }
    s += ";\n";
    return dup_string(s);
}

char* dipb_tree_newick(int n_nodes, int root_node, const int32_t* head, const int32_t* e, const int32_t* nxt,
                       const double* len, const char* const* names) {
    (void)n_nodes;
    std::string s;
    struct Frame { int node, from, slot; };
    std::vector<Frame> st;
    st.push_back({root_node, -1, -2});
    while (!st.empty()) {
        Frame& f = st.back();
        if (nxt[head[f.node]] == -1) { s += names[f.node]; st.pop_back(); continue; }
        int q;
        bool fresh = f.slot == -2;
        if (fresh) { s += '('; q = head[f.node]; }
        else { s += ':'; append_g(s, len[f.slot]); q = nxt[f.slot]; }
        while (q != -1 && e[q] == f.from) q = nxt[q];
        if (q == -1) { s += ')'; st.pop_back(); continue; }
        if (!fresh) s += ',';
        f.slot = q;
        int child = e[q], me = f.node;
        st.push_back({child, me, -2});
    }
    s += ";\n";
    return dup_string(s);
}

void dipb_free_str(char* s) { free(s); }

int dipb_format_g(double v, char* out32) {
    int n = fmt_g6_fast(v, out32);
    if (n == 0) n = snprintf(out32, 32, "%g", v);
    out32[n] = 0;
    return n;
}

int dipb_phylip_write(const char* path, int n, const double* D, const char* const* names, int lower) {
    if (!path || !D || !names || n < 1) { dipb::set_error("dipb_phylip_write: bad argument"); return DIPB_E_ARG; }
    FILE* f = fopen(path, "w");
    if (!f) { dipb::set_error("dipb_phylip_write: cannot open %s", path); return DIPB_E_ARG; }
    std::vector<char> buf((size_t)1 << 22);
    setvbuf(f, buf.data(), _IOFBF, buf.size());
    fprintf(f, "%d\n", n);
    std::string line;
    char num[40];
    for (int i = 0; i < n; i++) {
        line.assign(names[i]);
        const int cols = lower ? i : n;
        for (int j = 0; j < cols; j++) {
            const int k = snprintf(num, sizeof num, " %.9g", D[(size_t)i * n + j]);
            line.append(num, (size_t)k);
        }
        line.push_back('\n');
        if (fwrite(line.data(), 1, line.size(), f) != line.size()) { fclose(f); dipb::set_error("dipb_phylip_write: write failed"); return DIPB_E_ARG; }
    }
    if (fclose(f) != 0) { dipb::set_error("dipb_phylip_write: close failed"); return DIPB_E_ARG; }
    return 0;
}

int dipb_backbone_from_newick(const char* newick, int total_leaves, int32_t* head, int32_t* e, int32_t* nxt,
                              int32_t* belong, double* len, char** leaf_names_out) {
    if (!newick || total_leaves < 2 || !head || !e || !nxt || !belong || !len) {
        dipb::set_error("dipb_backbone_from_newick: bad argument");
        return DIPB_E_ARG;
    }
    struct N { int parent; int idx; double bl; std::vector<int> ch; std::string name; };
    std::vector<N> nodes;
    std::vector<int> stack;
    int next_internal = total_leaves, next_leaf = 0;
    int cur = -1;  // node whose label/length is being read
    const char* p = newick;
    int root = -1;
    while (*p && *p != ';') {
        char c = *p;
        if (c == '(') {
            N nd; nd.parent = stack.empty() ? -1 : stack.back(); nd.idx = next_internal++; nd.bl = 0;
            nodes.push_back(nd);
            int id = (int)nodes.size() - 1;
            if (nd.parent >= 0) nodes[nd.parent].ch.push_back(id); else root = id;
            stack.push_back(id);
            cur = -1; p++;
        } else if (c == ',') { cur = -1; p++; }
        else if (c == ')') {
            if (stack.empty()) { dipb::set_error("newick: unbalanced ')'"); return DIPB_E_ARG; }
            cur = stack.back(); stack.pop_back(); p++;
        } else if (c == ':') {
            p++;
            char* endp = nullptr;
            float f = strtof(p, &endp);  // src/tree.cpp:268,289 parse with stof
            if (endp == p) { dipb::set_error("newick: missing branch length"); return DIPB_E_ARG; }
            if (cur >= 0) nodes[cur].bl = (double)f;
            p = endp;
        } else if (c == ' ' || c == '\n' || c == '\r' || c == '\t') { p++; }
        else {
            std::string label;
            if (c == '\'') { p++; while (*p && *p != '\'') label += *p++; if (*p == '\'') p++; }
            else while (*p && !strchr(":,();", *p)) label += *p++;
            if (cur == -1) {
                if (stack.empty()) { dipb::set_error("newick: leaf outside parentheses"); return DIPB_E_ARG; }
                N nd; nd.parent = stack.back(); nd.idx = next_leaf++; nd.bl = 0; nd.name = label;
                nodes.push_back(nd);
                cur = (int)nodes.size() - 1;
                nodes[nd.parent].ch.push_back(cur);
            }  // labels on internal nodes are ignored, as in the reference
        }
    }
    if (!stack.empty() || root < 0) { dipb::set_error("newick: unbalanced parentheses"); return DIPB_E_ARG; }
    const int B = next_leaf;
    if (B > total_leaves || 2 * (size_t)(nodes.size() - 1) > 8 * (size_t)total_leaves) {
        dipb::set_error("newick: %d leaves do not fit total_leaves=%d", B, total_leaves);
        return DIPB_E_ARG;
    }
    for (int i = 0; i < 2 * total_leaves; i++) head[i] = -1;
    for (int i = 0; i < 8 * total_leaves; i++) { e[i] = -1; nxt[i] = -1; belong[i] = -1; len[i] = 2; }
    // post-order, two slots per non-root node: child->parent then parent->child
    struct F { int node; size_t k; };
    std::vector<F> st;
    st.push_back({root, 0});
    int edge = 0;
    while (!st.empty()) {
        F& f = st.back();
        if (f.k < nodes[f.node].ch.size()) { int c2 = nodes[f.node].ch[f.k++]; st.push_back({c2, 0}); continue; }
        int v = f.node;
        st.pop_back();
        if (nodes[v].parent < 0) continue;
        int x = nodes[v].idx, y = nodes[nodes[v].parent].idx;
        e[edge] = y; len[edge] = nodes[v].bl; belong[edge] = x; nxt[edge] = head[x]; head[x] = edge; edge++;
        e[edge] = x; len[edge] = nodes[v].bl; belong[edge] = y; nxt[edge] = head[y]; head[y] = edge; edge++;
    }
    // The placement kernels take the backbone as a rooted binary tree: 2B-2 edges = 4B-4 directed slots
    // (src/placement_close_k.cu:887 assumes it silently; a trifurcating root or unary nodes would leave unused
    // slots that the device code walks).
    if (edge != 4 * B - 4) {
        dipb::set_error("newick: backbone with %d leaves has %d edges; a rooted binary tree (%d edges) is required", B, edge / 2, 2 * B - 2);
        return DIPB_E_ARG;
    }
    if (leaf_names_out) {
        std::vector<std::string> nm(B);
        for (auto& nd : nodes) if (nd.ch.empty() && nd.idx < B) nm[nd.idx] = nd.name;
        std::string all;
        for (int i = 0; i < B; i++) { all += nm[i]; all += '\n'; }
        *leaf_names_out = dup_string(all);
    }
    return B;
}

}  // extern "C"
