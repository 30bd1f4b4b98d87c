// dipper -- command-line driver with the reference's contract (src/tree_generation.cu:33-99,
// 159-648): -i r|m|d, -I, -O, -o t, -m 0..3, -p, -k, -s, -d, -a/--add, -t, -h.
// Host C++ only; the GPU is reached through the C ABI of libdipper_b200.so.
// Deliberate differences: --device (reference hard-codes device 1),
// --seed / --no-shuffle pin the input permutation (reference seeds with time(NULL)),
// -p is honoured (the reference fills its placement mode from -m, src/tree_generation.cu:222-223, so there an explicit
// -m 0 on 30 000 .. 1 000 000 sequences runs the exact mode and -p is ignored; here -p 0 selects it, as documented),
// Mash + NJ sketches first (reference bug B3), Boost/TBB are not needed.
#include <getopt.h>
#include <zlib.h>

#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <ctime>
#include <fstream>
#include <functional>
#include <iostream>
#include <random>
#include <string>
#include <thread>
#include <unordered_map>
#include <vector>

#include "../../include/dipper_b200.h"
#include "../../include/dipper_host.h"

using Clock = std::chrono::high_resolution_clock;
static long ms_since(Clock::time_point t0) { return (long)std::chrono::duration_cast<std::chrono::milliseconds>(Clock::now() - t0).count(); }

static void usage() {
    std::cerr <<
        "DIPPER Command Line Arguments (dipper_b200)\n"
        "Required Options:\n"
        "  -i, --input-format arg   d - distance matrix in PHYLIP format\n"
        "                           r - unaligned sequences in FASTA format\n"
        "                           m - aligned sequences in FASTA format\n"
        "  -I, --input-file arg     Input file path (FASTA may be gzip-compressed)\n"
        "  -O, --output-file arg    Output file path\n"
        "Optional Options:\n"
        "  -o, --output-format arg  t - phylogenetic tree in Newick format (default)\n"
        "                           d - distance matrix in PHYLIP format (lower-triangular; -i r / -i m)\n"
        "  -m, --algorithm arg      0 - default mode (NJ < 30000 <= placement < 1000000 <= divide-and-conquer)\n"
        "                           1 - force placement, 2 - force conventional NJ, 3 - force divide-and-conquer\n"
        "  -p, --placement-mode arg 0 - exact mode, 1 - k-closest mode (default)\n"
        "      --devices a,b,..     several GPUs of this box for aligned input: NJ matrix row blocks / -m 3 queries\n"
        "  -k, --kmer-size arg      K-mer size, 2-32 (default: 15)\n"
        "  -s, --sketch-size arg    Sketch size (default: 1000)\n"
        "  -d, --distance-type arg  1 - uncorrected (default, as the reference ships), 2 - JC, 3 - Tajima-Nei,\n"
        "                           4 - K2P, 5 - Tamura, 6 - Jin-Nei\n"
        "  -a, --add                Add query sequences to a backbone tree using k-closest placement\n"
        "  -t, --input-tree arg     Input backbone tree (Newick), required with --add\n"
        "      --device arg         CUDA device ordinal (default 0)\n"
        "      --seed arg           seed of the input-order shuffle (default: time, like the reference)\n"
        "      --no-shuffle         keep the input order\n"
        "  -h, --help               Print this help message\n";
}

static bool read_fasta(const std::string& path, std::vector<std::string>& seqs, std::vector<std::string>& names) {
    gzFile f = gzopen(path.c_str(), "r");
    if (!f) return false;
    gzbuffer(f, 1 << 20);
    std::vector<char> buf(1 << 20);
    std::string cur;
    bool have = false;
    while (gzgets(f, buf.data(), (int)buf.size())) {
        size_t len = strlen(buf.data());
        bool eol = len && buf[len - 1] == '\n';
        while (len && (buf[len - 1] == '\n' || buf[len - 1] == '\r')) len--;
        if (buf[0] == '>') {
            if (have) seqs.push_back(std::move(cur));
            cur.clear();
            std::string nm(buf.data() + 1, len ? len - 1 : 0);
            size_t sp = nm.find_first_of(" \t");   // kseq: name = up to first whitespace
            if (sp != std::string::npos) nm.resize(sp);
            names.push_back(nm);
            have = true;
        } else if (have) {
            cur.append(buf.data(), len);
        }
        (void)eol;
    }
    if (have) seqs.push_back(std::move(cur));
    gzclose(f);
    return true;
}

#define CHECK(call)                                                                   \
    do {                                                                              \
        int rc__ = (call);                                                            \
        if (rc__ != 0) {                                                              \
            std::cerr << "dipper: " << #call << " failed: " << dipb_last_error() << "\n"; \
            return 1;                                                                 \
        }                                                                             \
    } while (0)

static void parallel_for(size_t n, const std::function<void(size_t)>& fn) {
    unsigned nt = std::max(1u, std::thread::hardware_concurrency());
    if (n < 64) nt = 1;
    std::vector<std::thread> th;
    for (unsigned t = 0; t < nt; t++)
        th.emplace_back([&, t]() { for (size_t i = t; i < n; i += nt) fn(i); });
    for (auto& x : th) x.join();
}

static int write_nj(dipb_matrix* M, const std::vector<std::string>& names, std::ofstream& out) {
    int n = dipb_matrix_n(M);
    std::vector<int32_t> c0(n), c1(n);
    std::vector<double> l0(n), l1(n);
    CHECK(dipb_nj(M, DIPB_NJ_AUTO, c0.data(), c1.data(), l0.data(), l1.data()));
    std::vector<const char*> nm(n);
    for (int i = 0; i < n; i++) nm[i] = names[i].c_str();
    char* s = dipb_nj_newick(n, c0.data(), c1.data(), l0.data(), l1.data(), nm.data());
    out << s;
    dipb_free_str(s);
    return 0;
}

static int write_tree(dipb_tree* T, const std::vector<std::string>& names, std::ofstream& out) {
    int n = dipb_tree_n(T);
    std::vector<int32_t> head(2 * (size_t)n), e(8 * (size_t)n), nxt(8 * (size_t)n), belong(8 * (size_t)n);
    std::vector<double> len(8 * (size_t)n);
    CHECK(dipb_tree_export(T, head.data(), e.data(), nxt.data(), belong.data(), len.data()));
    std::vector<const char*> nm(2 * (size_t)n, "");
    for (size_t i = 0; i < names.size() && i < (size_t)n; i++) nm[i] = names[i].c_str();
    char* s = dipb_tree_newick(2 * n, n, head.data(), e.data(), nxt.data(), len.data(), nm.data());
    out << s;
    dipb_free_str(s);
    return 0;
}

int main(int argc, char** argv) {
    auto t_input = Clock::now();
    std::string in = "r", out_fmt = "t", algo = "0", placemode = "1", input, output, tree_file;
    long k = 15, sketch = 1000, dist_type = 1, device = 0;
    bool add = false, shuffle = true, have_seed = false, help = false;
    std::vector<int> devices;
    unsigned long seed = 0;
    static option opts[] = {{"input-format", 1, 0, 'i'}, {"input-file", 1, 0, 'I'}, {"output-file", 1, 0, 'O'},
                            {"output-format", 1, 0, 'o'}, {"algorithm", 1, 0, 'm'}, {"placement-mode", 1, 0, 'p'},
                            {"kmer-size", 1, 0, 'k'}, {"sketch-size", 1, 0, 's'}, {"distance-type", 1, 0, 'd'},
                            {"add", 0, 0, 'a'}, {"input-tree", 1, 0, 't'}, {"help", 0, 0, 'h'},
                            {"device", 1, 0, 1000}, {"seed", 1, 0, 1001}, {"no-shuffle", 0, 0, 1002}, {"devices", 1, 0, 1003}, {0, 0, 0, 0}};
    int c;
    auto to_long = [](const char* s, long dflt) { char* e; long v = strtol(s, &e, 10); return (e == s) ? dflt : v; };
    while ((c = getopt_long(argc, argv, "i:I:O:o:m:p:k:s:d:at:h", opts, nullptr)) != -1) {
        switch (c) {
            case 'i': in = optarg; break;
            case 'I': input = optarg; break;
            case 'O': output = optarg; break;
            case 'o': out_fmt = optarg; break;
            case 'm': algo = optarg; break;
            case 'p': placemode = optarg; break;
            case 'k': k = to_long(optarg, 15); break;          // bad values fall back to defaults (:192-208)
            case 's': sketch = to_long(optarg, 1000); break;
            case 'd': dist_type = to_long(optarg, 1); break;
            case 'a': add = true; break;
            case 't': tree_file = optarg; break;
            case 'h': help = true; break;
            case 1000: device = to_long(optarg, 0); break;
            case 1001: seed = strtoul(optarg, nullptr, 10); have_seed = true; break;
            case 1002: shuffle = false; break;
            case 1003: {   // --devices 0,1,2,3: several GPUs of this box (aligned input: NJ matrix and -m 3)
                for (const char* q = optarg; *q;) { char* e2; long v = strtol(q, &e2, 10); if (e2 == q) break; devices.push_back((int)v); q = *e2 == ',' ? e2 + 1 : e2; }
                break;
            }
            default: usage(); return 1;
        }
    }
    if (help) { usage(); return 0; }
    if (input.empty() || output.empty()) {
        std::cerr << "\033[31mthe options '--input-file' and '--output-file' are required\033[0m\n";
        usage();
        return 1;
    }
    if (add && tree_file.empty()) {
        std::cerr << "\033[31mBackbone tree (--input-tree/-t) is required with --add option\033[0m\n";
        usage();
        return 1;
    }
    if ((out_fmt != "t" && out_fmt != "d") || (in != "r" && in != "m" && in != "d")) { printf("Invalid input-output combinations!!!!!\n"); return 1; }
    if (out_fmt == "d" && (in == "d" || add)) { std::cerr << "-o d needs sequence input (-i r or -i m) and no --add\n"; return 1; }
    std::ofstream output_(output.c_str());
    if (!output_) { std::cerr << "ERROR: cant open output file: " << output << "\n"; return 1; }

    if (!devices.empty()) device = devices[0];
    dipb_ctx* ctx = nullptr;
    if (dipb_init((int)device, &ctx) != 0) { std::cerr << "Failed to set CUDA device: " << dipb_last_error() << std::endl; return -1; }
    const int placement_thr = 30000, dc_thr = 1000000;

    // ------------------------------------------------------------------ -i d
    if (in == "d") {
        if (add) { std::cerr << "Adding new sequnces only supported with input aligned and unaligned sequences\n"; return 1; }
        FILE* f = fopen(input.c_str(), "r");
        if (!f) { std::cerr << "Cannot open file: " << input << std::endl; return 1; }
        int n = 0;
        if (fscanf(f, "%d", &n) != 1 || n < 2) { std::cerr << "Bad PHYLIP header\n"; return 1; }
        std::vector<std::string> names(n);
        std::vector<double> tri((size_t)n * (n - 1) / 2);
        std::vector<char> tok(256);
        for (int i = 0; i < n; i++) {
            char nm[4096];
            if (fscanf(f, "%4095s", nm) != 1) { std::cerr << "Bad PHYLIP row " << i << "\n"; return 1; }
            names[i] = nm;
            for (int j = 0; j < i; j++) {
                char num[128];
                if (fscanf(f, "%127s", num) != 1) { std::cerr << "Bad PHYLIP row " << i << "\n"; return 1; }
                tri[(size_t)i * (i - 1) / 2 + j] = (double)strtof(num, nullptr);   // stof, src/matrix_reader.cu:42
            }
            int ch;   // skip the rest of a full-matrix row
            while ((ch = fgetc(f)) != '\n' && ch != EOF) {}
        }
        fclose(f);
        dipb_matrix* M = nullptr;
        CHECK(dipb_matrix_from_host(ctx, tri.data(), n, 0, &M));
        bool place = algo == "1" || (algo == "0" && n >= placement_thr && n < dc_thr);
        if (place) {
            std::cerr << (placemode == "0" ? "Using exact placement mode\n" : "Using k-closest placement mode\n");
            dipb_dist_source src{};
            src.matrix = M;
            dipb_tree* T = nullptr;
            if (placemode == "0") CHECK(dipb_place_exact(ctx, &src, n, &T));
            else CHECK(dipb_place_kclosest(ctx, &src, n, &T));
            if (write_tree(T, names, output_)) return 1;
            dipb_tree_free(T);
        } else if (algo == "3" || (algo == "0" && n >= dc_thr)) {
            std::cerr << "Divide-and-conquer mode not supported with input matrix\n";
            return 1;
        } else {
            std::cerr << "Using conventional NJ\n";
            if (write_nj(M, names, output_)) return 1;
        }
        dipb_matrix_free(M);
        dipb_destroy(ctx);
        return 0;
    }

    // ------------------------------------------------------------------ FASTA inputs
    // Plain FASTA goes through the parallel memory-mapped reader, which also packs (dipb_fasta_open); gzip input keeps the
    // zlib line reader and is packed below.
    const bool aligned = in == "m";
    std::vector<std::string> seqs, names_in;
    dipb_fasta* fa = nullptr;
    {
        int frc = dipb_fasta_open(input.c_str(), aligned ? 4 : 2, 0, &fa);
        if (frc == DIPB_E_UNSUPPORTED) {
            fa = nullptr;
            if (!read_fasta(input, seqs, names_in)) { fprintf(stderr, "ERROR: cant open file: %s\n", input.c_str()); return 1; }
        } else if (frc != 0) {
            fprintf(stderr, "ERROR: cant open file: %s\n", input.c_str());
            return 1;
        } else {
            names_in.resize(dipb_fasta_count(fa));
            for (size_t i = 0; i < names_in.size(); i++) names_in[i] = dipb_fasta_name(fa, i);
        }
    }
    const size_t n = names_in.size();
    if (n < 2) { std::cerr << "ERROR: need at least two sequences\n"; return 1; }
    std::vector<std::string> names(n);
    std::vector<size_t> ids(n);          // ids[i] = row of input sequence i
    int backbone = 0;
    std::vector<int32_t> bb_head, bb_e, bb_nxt, bb_belong;
    std::vector<double> bb_len;
    if (add) {
        std::cerr << "Read " << n << " sequences from input file.\n";
        std::ifstream tf(tree_file);
        if (!tf) { std::cerr << "ERROR: Unable to open input tree file: " << tree_file << "\n"; return 1; }
        std::string nwk;
        std::getline(tf, nwk);
        bb_head.resize(2 * n); bb_e.resize(8 * n); bb_nxt.resize(8 * n); bb_belong.resize(8 * n); bb_len.resize(8 * n);
        char* leafs = nullptr;
        backbone = dipb_backbone_from_newick(nwk.c_str(), (int)n, bb_head.data(), bb_e.data(), bb_nxt.data(), bb_belong.data(), bb_len.data(), &leafs);
        if (backbone < 0) { std::cerr << "dipper: " << dipb_last_error() << "\n"; return 1; }
        std::unordered_map<std::string, int> leaf_idx;
        {
            std::string all(leafs);
            dipb_free_str(leafs);
            size_t pos = 0; int q = 0;
            while (pos < all.size()) { size_t nl = all.find('\n', pos); leaf_idx[all.substr(pos, nl - pos)] = q++; pos = nl + 1; }
        }
        std::cerr << "Tree loaded successfully with " << backbone << " leaves.\n";
        size_t next = backbone;   // idMap, src/tree_generation.cu:271-282
        for (size_t i = 0; i < n; i++) {
            auto it = leaf_idx.find(names_in[i]);
            ids[i] = it == leaf_idx.end() ? next++ : (size_t)it->second;
        }
        if (next != n) { std::cerr << "ERROR: " << (n - next) << " backbone tips have no sequence in the input file\n"; return 1; }
    } else {
        for (size_t i = 0; i < n; i++) ids[i] = i;
        if (shuffle) {   // :341-344
            std::mt19937 rnd(have_seed ? seed : (unsigned long)time(NULL));
            std::shuffle(ids.begin(), ids.end(), rnd);
        }
    }
    std::vector<std::vector<uint64_t>> packed(fa ? 0 : n);
    std::vector<uint64_t> lens(n);
    std::vector<const uint64_t*> ptrs(n);
    if (fa) {
        const uint64_t *fl = dipb_fasta_lengths(fa), *fo = dipb_fasta_word_offsets(fa), *fw = dipb_fasta_words(fa);
        for (size_t i = 0; i < n; i++) { ptrs[ids[i]] = fw + fo[i]; lens[ids[i]] = fl[i]; names[ids[i]] = names_in[i]; }
    } else {
        parallel_for(n, [&](size_t i) {
            const std::string& s = seqs[i];
            std::vector<uint64_t> w((s.size() + (aligned ? 15 : 31)) / (aligned ? 16 : 32));
            if (aligned) dipb_pack4(s.data(), s.size(), w.data()); else dipb_pack2(s.data(), s.size(), w.data());
            packed[ids[i]] = std::move(w);
            lens[ids[i]] = s.size();
            names[ids[i]] = names_in[i];
        });
        for (size_t i = 0; i < n; i++) ptrs[i] = packed[i].data();
    }
    std::cerr << "Input in: " << ms_since(t_input) << " ms\n";

    if (devices.size() > 1 && aligned && !add && out_fmt == "t" &&
        (algo == "2" || algo == "3" || (algo == "0" && (n < (size_t)placement_thr || n >= (size_t)dc_thr)))) {
        // several GPUs, one process (csrc/multi.cu): distance row blocks or D&C queries sharded, tree on devices[0]
        for (size_t i = 1; i < n; i++) if (lens[i] != lens[0]) { std::cerr << "dipper: aligned input requires equal sequence lengths\n"; return 1; }
        const size_t comp = (lens[0] + 15) / 16;
        std::vector<uint64_t> flat(n * comp);
        parallel_for(n, [&](size_t i) { memcpy(flat.data() + i * comp, ptrs[i], comp * sizeof(uint64_t)); });
        dipb_multi* md = nullptr;
        CHECK(dipb_multi_init(devices.data(), (int)devices.size(), &md));
        CHECK(dipb_multi_msa_upload_flat(md, flat.data(), n, lens[0]));
        auto t_tree2 = Clock::now();
        if (algo == "3" || (algo == "0" && n >= (size_t)dc_thr)) {
            std::cerr << "Using divide-and-conquer mode on " << devices.size() << " devices\n";
            dipb_tree* T = nullptr;
            CHECK(dipb_multi_dc(md, (int)dist_type, (int)(n / 20), &T));
            if (write_tree(T, names, output_)) return 1;
            dipb_tree_free(T);
        } else {
            std::cerr << "Using conventional NJ, distance matrix on " << devices.size() << " devices\n";
            dipb_matrix* M = nullptr;
            CHECK(dipb_multi_msa_dist_matrix(md, (int)dist_type, &M));
            if (write_nj(M, names, output_)) return 1;
            dipb_matrix_free(M);
        }
        std::cerr << "Tree Created in: " << ms_since(t_tree2) << " ms\n";
        dipb_multi_destroy(md);
        if (fa) dipb_fasta_close(fa);
        dipb_destroy(ctx);
        return 0;
    }
    auto t_alloc = Clock::now();
    dipb_msa* msa = nullptr;
    dipb_mash* mash = nullptr;
    dipb_dist_source src{};
    src.dist_type = (int)dist_type;
    if (aligned) {
        CHECK(dipb_msa_upload(ctx, ptrs.data(), lens.data(), n, &msa));
        src.msa = msa;
    } else {
        CHECK(dipb_mash_upload(ctx, ptrs.data(), lens.data(), n, (int)k, (int)sketch, &mash));
        std::cerr << "Allocated in: " << ms_since(t_alloc) << " ms\n";
        auto t_sk = Clock::now();
        CHECK(dipb_mash_sketch(mash));
        std::cerr << "Sketch Created in: " << ms_since(t_sk) << " ms\n";
        src.mash = mash;
    }
    if (aligned) std::cerr << "Allocated in: " << ms_since(t_alloc) << " ms\n";

    auto t_tree = Clock::now();
    if (out_fmt == "d") {
        // -o d (documented by the reference, docs/index.md:114, not implemented there): lower-triangular PHYLIP
        output_.close();
        dipb_matrix* M = nullptr;
        if (aligned) CHECK(dipb_msa_dist_matrix(msa, (int)dist_type, &M));
        else CHECK(dipb_mash_dist_matrix(mash, &M));
        std::vector<double> D(n * n);
        CHECK(dipb_matrix_to_host(M, D.data()));
        dipb_matrix_free(M);
        std::vector<const char*> nm(n);
        for (size_t i = 0; i < n; i++) nm[i] = names[i].c_str();
        CHECK(dipb_phylip_write(output.c_str(), (int)n, D.data(), nm.data(), 1));
        std::cerr << "Distance matrix written in: " << ms_since(t_tree) << " ms\n";
    } else if (add) {
        dipb_tree* T = nullptr;
        CHECK(dipb_place_add(ctx, &src, (int)n, backbone, bb_head.data(), bb_e.data(), bb_nxt.data(), bb_belong.data(), bb_len.data(), &T));
        if (write_tree(T, names, output_)) return 1;
        dipb_tree_free(T);
    } else if (algo == "1" || (algo == "0" && n >= (size_t)placement_thr && n < (size_t)dc_thr)) {
        std::cerr << "Using " << (placemode == "0" ? "exact placement mode\n" : "k-closest placement mode\n");
        dipb_tree* T = nullptr;
        if (placemode == "0") CHECK(dipb_place_exact(ctx, &src, (int)n, &T));
        else CHECK(dipb_place_kclosest(ctx, &src, (int)n, &T));
        if (write_tree(T, names, output_)) return 1;
        dipb_tree_free(T);
    } else if (algo == "3" || (algo == "0" && n >= (size_t)dc_thr)) {
        std::cerr << "Using divide-and-conquer mode\n";
        dipb_tree* T = nullptr;
        CHECK(dipb_dc(ctx, &src, (int)n, (int)(n / 20), &T));
        if (write_tree(T, names, output_)) return 1;
        dipb_tree_free(T);
    } else {
        std::cerr << "Using conventional NJ\n";
        if (n >= 40000) std::cerr << "Warning: forcing conventional NJ on large datasets might result in unexpected behavior\n";
        dipb_matrix* M = nullptr;
        if (aligned) CHECK(dipb_msa_dist_matrix(msa, (int)dist_type, &M));
        else CHECK(dipb_mash_dist_matrix(mash, &M));
        if (write_nj(M, names, output_)) return 1;
        dipb_matrix_free(M);
    }
    std::cerr << "Tree Created in: " << ms_since(t_tree) << " ms\n";
    if (msa) dipb_msa_free(msa);
    if (mash) dipb_mash_free(mash);
    if (fa) dipb_fasta_close(fa);
    dipb_destroy(ctx);
    return 0;
}
