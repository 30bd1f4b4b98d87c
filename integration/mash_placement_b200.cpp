// mash_placement_b200.cpp -- the drop-in shim: the struct API of the reference's UNMODIFIED src/mash_placement.cuh,
// implemented on libdipper_b200.so (include/dipper_b200.h, include/dipper_host.h).
//
// A DIPPER maintainer adds this one file, links -ldipper_b200 and drops the nine kernel files
// (src/MSA.cu, mash.cu, neighborJoining.cu, placement.cu, placement_close_k.cu, matrix_reader.cu stays,
// src/divide_and_conquer/{msa,mash,placement_close_k}.cu) from CMakeLists.txt:22-37; main (src/tree_generation.cu) and
// everything above it stay as they are.  The file is COMPILED AND TESTED here: oracle/build_ref.sh links it with
// oracle/ref_driver.cu (the same main that drives the reference's own objects) into oracle/_ref/dipper_ref_b200, and
// tests/test_shim_gpu.py diffs the outputs of the two binaries mode by mode.
//
// Every method cites the reference definition it replaces (paths relative to the reference root).  Like the reference,
// errors print "Gpu_ERROR ..." and exit(1); the public device-pointer fields the reference exposes (d_head, d_e, ...)
// are filled so that callers that read them keep working.
#include "mash_placement.cuh"

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <functional>
#include <vector>

#include "../include/dipper_b200.h"
#include "../include/dipper_host.h"

namespace {
dipb_ctx* g_ctx = nullptr;      // one context: the reference runs on one device (src/tree_generation.cu:240)
dipb_msa* g_msa = nullptr;
dipb_mash* g_mash = nullptr;
dipb_matrix* g_mat = nullptr;
dipb_tree* g_tree = nullptr;
dipb_dc_state* g_dc = nullptr;
std::vector<int32_t> g_bb_head, g_bb_e, g_bb_nxt, g_bb_belong;   // backbone loaded by initializeDeviceArrays(Tree*)
std::vector<double> g_bb_len;
int g_backbone = 0;

[[noreturn]] void die(const char* what) {
    fprintf(stderr, "Gpu_ERROR: %s: %s\n", what, dipb_last_error());
    exit(1);
}
dipb_ctx* ctx() {
    if (!g_ctx) {
        int dev = 0;
        if (const char* e = getenv("DIPPER_DEVICE")) dev = atoi(e);   // the reference hard-codes device 1
        if (dipb_init(dev, &g_ctx)) die("dipb_init");
    }
    return g_ctx;
}
dipb_dist_source source(const MashPlacement::Param& p) {
    dipb_dist_source s{};
    s.dist_type = (int)p.distanceType;
    if (p.in == "m") s.msa = g_msa;
    else if (p.in == "r") s.mash = g_mash;
    else s.matrix = g_mat;
    return s;
}
template <class T>
void publish_tree(T* self, dipb_tree* t) {   // the reference's public device-pointer fields
    int32_t *head, *e, *nxt, *belong, *cid;
    double *len, *cdis;
    if (dipb_tree_device_arrays(t, &head, &e, &nxt, &belong, &len, &cid, &cdis)) die("dipb_tree_device_arrays");
    self->d_head = head; self->d_e = e; self->d_nxt = nxt; self->d_belong = belong; self->d_len = len;
    self->d_closest_id = cid; self->d_closest_dis = cdis;
}
void print_tree(dipb_tree* t, int n, const std::vector<std::string>& name, std::ofstream& out) {
    std::vector<int32_t> head(2 * (size_t)n), e(8 * (size_t)n), nxt(8 * (size_t)n), belong(8 * (size_t)n);
    std::vector<double> len(8 * (size_t)n);
    if (dipb_tree_export(t, head.data(), e.data(), nxt.data(), belong.data(), len.data())) die("printTree");
    std::vector<const char*> nm(2 * (size_t)n, "");
    for (size_t i = 0; i < name.size() && i < (size_t)n; i++) nm[i] = name[i].c_str();
    char* s = dipb_tree_newick(2 * n, n, head.data(), e.data(), nxt.data(), len.data(), nm.data());
    out << s;
    dipb_free_str(s);
}
}  // namespace

namespace MashPlacement {

// ---- aligned input: src/mash_placement.cuh:89-99, src/MSA.cu:14-72, 271-282
void MSADeviceArrays::allocateDeviceArrays(uint64_t** h_compressedSeqs, uint64_t* h_seqLengths, size_t num, Param&) {
    numSequences = num;
    seqLen = (int)h_seqLengths[0];
    if (dipb_msa_upload(ctx(), h_compressedSeqs, h_seqLengths, num, &g_msa)) die("MSADeviceArrays::allocateDeviceArrays");
}
void MSADeviceArrays::deallocateDeviceArrays() { dipb_msa_free(g_msa); g_msa = nullptr; }
void MSADeviceArrays::distConstructionOnGpu(Param& p, int rowId, double* d_mashDist) const {
    if (dipb_msa_dist_row(g_msa, (int)p.distanceType, rowId, d_mashDist)) die("MSADeviceArrays::distConstructionOnGpu");
}

// ---- unaligned input: src/mash_placement.cuh:34-50, src/mash.cu:14-122, 386-424, 457-471
void MashDeviceArrays::allocateDeviceArrays(uint64_t** h_compressedSeqs, uint64_t* h_seqLengths, size_t num, Param& p) {
    numSequences = num;
    if (dipb_mash_upload(ctx(), h_compressedSeqs, h_seqLengths, num, (int)p.kmerSize, (int)p.sketchSize, &g_mash)) die("MashDeviceArrays::allocateDeviceArrays");
}
void MashDeviceArrays::sketchConstructionOnGpu(Param& p) {
    if (dipb_mash_sketch(g_mash)) die("MashDeviceArrays::sketchConstructionOnGpu");
    // the reference leaves the sketches on the host too, transposed: h_hashList[t * n + seq] (src/mash.cu:381-383,418-419)
    const size_t n = numSequences, s = p.sketchSize;
    std::vector<uint64_t> sk(n * s);
    if (dipb_mash_get_sketches(g_mash, sk.data())) die("dipb_mash_get_sketches");
    h_hashList = new uint64_t[n * s];
    for (size_t q = 0; q < n; q++)
        for (size_t t = 0; t < s; t++) h_hashList[t * n + q] = sk[q * s + t];
}
void MashDeviceArrays::deallocateDeviceArrays() { dipb_mash_free(g_mash); g_mash = nullptr; }
void MashDeviceArrays::distConstructionOnGpu(Param&, int rowId, double* d_mashDist) const {
    if (dipb_mash_dist_row(g_mash, rowId, d_mashDist)) die("MashDeviceArrays::distConstructionOnGpu");
}

// ---- conventional NJ: src/mash_placement.cuh:199-214, src/neighborJoining.cu:35-85, 197-271
void NJDeviceArrays::getDismatrix(int numSequences, Param& p, const MashDeviceArrays&, MatrixReader& rd, const MSADeviceArrays&) {
    d_numSequences = numSequences;
    int rc;
    if (p.in == "m") rc = dipb_msa_dist_matrix(g_msa, (int)p.distanceType, &g_mat);
    else if (p.in == "r") rc = dipb_mash_dist_matrix(g_mash, &g_mat);
    else {
        // -i d: rows come from the reference's own MatrixReader (src/matrix_reader.cu:23-44), one device row at a time
        std::vector<double> tri((size_t)numSequences * (numSequences - 1) / 2);
        double* d_row = nullptr;
        cudaMalloc(&d_row, sizeof(double) * numSequences);
        for (int i = 1; i < numSequences; i++) {
            rd.distConstructionOnGpu(p, i, d_row);
            cudaMemcpy(tri.data() + (size_t)i * (i - 1) / 2, d_row, sizeof(double) * i, cudaMemcpyDeviceToHost);
        }
        cudaFree(d_row);
        rc = dipb_matrix_from_host(ctx(), tri.data(), numSequences, 0, &g_mat);
    }
    if (rc) die("NJDeviceArrays::getDismatrix");
    d_mashDist = dipb_matrix_device_ptr(g_mat);
}
void NJDeviceArrays::findNeighbourJoiningTree(std::vector<std::string>& name, std::ofstream& output_) {
    const int n = d_numSequences;
    std::vector<int32_t> c0(n), c1(n);
    std::vector<double> l0(n), l1(n);
    if (dipb_nj(g_mat, DIPB_NJ_AUTO, c0.data(), c1.data(), l0.data(), l1.data())) die("NJDeviceArrays::findNeighbourJoiningTree");
    std::vector<const char*> nm(n);
    for (int i = 0; i < n; i++) nm[i] = name[i].c_str();
    char* s = dipb_nj_newick(n, c0.data(), c1.data(), l0.data(), l1.data(), nm.data());   // same text as :252-270
    output_ << s;
    dipb_free_str(s);
}
void NJDeviceArrays::deallocateDeviceArrays() { dipb_matrix_free(g_mat); g_mat = nullptr; }

// ---- k-closest placement and add-tips: src/mash_placement.cuh:167-197, src/placement_close_k.cu:15-68, 126-264, 646-990, 568-643
void KPlacementDeviceArrays::allocateDeviceArrays(size_t num, int backbone) {
    numSequences = (int)num;
    backboneSize = backbone;
    bd = 2; idx = 0;
}
void KPlacementDeviceArrays::deallocateDeviceArrays() { dipb_tree_free(g_tree); g_tree = nullptr; }
void KPlacementDeviceArrays::findPlacementTree(Param& p, const MashDeviceArrays&, MatrixReader&, const MSADeviceArrays&) {
    dipb_dist_source s = source(p);
    if (dipb_place_kclosest(ctx(), &s, numSequences, &g_tree)) die("KPlacementDeviceArrays::findPlacementTree");
    publish_tree(this, g_tree);
}
void KPlacementDeviceArrays::initializeDeviceArrays(Tree* t) {
    // the host half of src/placement_close_k.cu:126-183: post-order dfs, per non-root node the slots child->parent then
    // parent->child; the closest-leaf lists (the device half, :241-260) are built by dipb_place_add
    const size_t N = (size_t)numSequences;
    g_bb_head.assign(2 * N, -1); g_bb_e.assign(8 * N, -1); g_bb_nxt.assign(8 * N, -1); g_bb_belong.assign(8 * N, -1);
    g_bb_len.assign(8 * N, 2.0);
    size_t edge = 0;
    std::function<void(Node*)> dfs = [&](Node* node) {
        for (Node* c : node->children) dfs(c);
        if (node->parent == nullptr) return;
        const int x = node->idx, y = node->parent->idx;
        g_bb_e[edge] = y; g_bb_len[edge] = node->bl; g_bb_belong[edge] = x; g_bb_nxt[edge] = g_bb_head[x]; g_bb_head[x] = (int)edge; edge++;
        g_bb_e[edge] = x; g_bb_len[edge] = node->bl; g_bb_belong[edge] = y; g_bb_nxt[edge] = g_bb_head[y]; g_bb_head[y] = (int)edge; edge++;
    };
    dfs(t->root);
    g_backbone = (int)t->m_numLeaves;
}
void KPlacementDeviceArrays::addQuery(Param& p, const MashDeviceArrays&, MatrixReader&, const MSADeviceArrays&) {
    dipb_dist_source s = source(p);
    if (dipb_place_add(ctx(), &s, numSequences, g_backbone, g_bb_head.data(), g_bb_e.data(), g_bb_nxt.data(), g_bb_belong.data(),
                       g_bb_len.data(), &g_tree)) die("KPlacementDeviceArrays::addQuery");
    publish_tree(this, g_tree);
}
void KPlacementDeviceArrays::printTree(std::vector<std::string> name, std::ofstream& output_) { print_tree(g_tree, numSequences, name, output_); }

// ---- exact placement mode: src/mash_placement.cuh:137-165, src/placement.cu:28-116, 505-789
void PlacementDeviceArrays::allocateDeviceArrays(size_t num) { numSequences = (int)num; bd = 2; idx = 0; }
void PlacementDeviceArrays::deallocateDeviceArrays() { dipb_tree_free(g_tree); g_tree = nullptr; }
void PlacementDeviceArrays::findPlacementTree(Param& p, const MashDeviceArrays&, MatrixReader&, const MSADeviceArrays&) {
    dipb_dist_source s = source(p);
    if (dipb_place_exact(ctx(), &s, numSequences, &g_tree)) die("PlacementDeviceArrays::findPlacementTree");
}
void PlacementDeviceArrays::printTree(std::vector<std::string> name, std::ofstream& output_) { print_tree(g_tree, numSequences, name, output_); }

// ---- divide and conquer: src/mash_placement.cuh:52-87,101-121,216-285, src/divide_and_conquer/{msa,mash,placement_close_k}.cu
void MSADeviceArraysDC::allocateDeviceArraysDC(uint64_t** h_compressedSeqs, uint64_t* h_seqLengths, size_t num, Param& p) {
    totalNumSequences = num;
    backboneSize = p.backboneSize;
    d_seqLen = (int)h_seqLengths[0];
    if (dipb_msa_upload(ctx(), h_compressedSeqs, h_seqLengths, num, &g_msa)) die("MSADeviceArraysDC::allocateDeviceArraysDC");
}
void MSADeviceArraysDC::deallocateDeviceArraysDC() { dipb_msa_free(g_msa); g_msa = nullptr; }
void MSADeviceArraysDC::distConstructionOnGpuForBackboneDC(Param& p, int rowId, double* d_mashDist) const {
    if (dipb_msa_dist_row(g_msa, (int)p.distanceType, rowId, d_mashDist)) die("MSADeviceArraysDC::distConstructionOnGpuForBackboneDC");
}
void MSADeviceArraysDC::distConstructionOnGpuDC(Param& p, int rowId, double* d_mashDist) const { distConstructionOnGpuForBackboneDC(p, rowId, d_mashDist); }
void MashDeviceArraysDC::allocateDeviceArraysDC(uint64_t** h_compressedSeqs, uint64_t* h_seqLengths, size_t num, Param& p) {
    totalNumSequences = num;
    backboneSize = p.backboneSize;
    if (dipb_mash_upload(ctx(), h_compressedSeqs, h_seqLengths, num, (int)p.kmerSize, (int)p.sketchSize, &g_mash)) die("MashDeviceArraysDC::allocateDeviceArraysDC");
}
void MashDeviceArraysDC::sketchConstructionOnGpuDC(Param&, uint64_t**, uint64_t*, uint64_t) {
    if (dipb_mash_sketch(g_mash)) die("MashDeviceArraysDC::sketchConstructionOnGpuDC");   // all sketches stay on the device
}
void MashDeviceArraysDC::deallocateDeviceArraysDC() { dipb_mash_free(g_mash); g_mash = nullptr; }

void KPlacementDeviceArraysDC::allocateDeviceArraysDC(size_t num, size_t totalNum) {
    numSequences = (int)num;
    totalNumSequences = (int)totalNum;
    bd = 2; idx = 0;
}
void KPlacementDeviceArraysDC::deallocateDeviceArraysDC() { dipb_tree_free(g_tree); g_tree = nullptr; }
void KPlacementDeviceArraysDC::findBackboneTreeDC(Param& p, const MashDeviceArraysDC&, MatrixReader&, const MSADeviceArraysDC&, const KPlacementDeviceArraysHostDC&) {
    dipb_dist_source s = source(p);                      // stage 1 (:731-935)
    if (dipb_dc_begin(ctx(), &s, totalNumSequences, numSequences, &g_dc)) die("findBackboneTreeDC");
}
void KPlacementDeviceArraysDC::findClustersDC(Param&, const MashDeviceArraysDC&, MatrixReader&, const MSADeviceArraysDC&, KPlacementDeviceArraysHostDC& host) {
    const int n = totalNumSequences, B = numSequences;   // stage 2 (:937-1113)
    host.clusterID = new int[n];
    for (int i = 0; i < B; i++) host.clusterID[i] = -1;
    if (dipb_dc_assign(g_dc, B, n, host.clusterID + B)) die("findClustersDC");
}
void KPlacementDeviceArraysDC::findClusterTreeDC(Param&, MashDeviceArraysDC&, MatrixReader&, MSADeviceArraysDC&, KPlacementDeviceArraysHostDC& host) {
    int ncl = 0;                                         // stage 3 (:1251-1535)
    if (dipb_dc_set_clusters(g_dc, host.clusterID, &ncl)) die("findClusterTreeDC (clusters)");
    if (dipb_dc_run_clusters(g_dc, 0, ncl)) die("findClusterTreeDC");
    if (dipb_dc_finish(g_dc, &g_tree)) die("findClusterTreeDC (finish)");
    g_dc = nullptr;
    publish_tree(this, g_tree);
}
void KPlacementDeviceArraysDC::printTreeDC(std::vector<std::string> name, std::ofstream& output_) { print_tree(g_tree, totalNumSequences, name, output_); }

}  // namespace MashPlacement
