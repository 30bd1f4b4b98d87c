// ref_driver.cu -- replacement main() for the reference's own CUDA objects.
// TEST / BASELINE INFRASTRUCTURE ONLY (lives under oracle/, built into oracle/_ref/).
//
// src/tree_generation.cu cannot be built here (needs Boost.program_options and TBB,
// neither installed, no network), so this file drives the reference's UNMODIFIED
// kernel files (compiled in place from /root/reference by oracle/build_ref.sh) through
// the struct API of src/mash_placement.cuh.  Documented deviations from the
// reference's main: device 0 instead of the hard-coded device 1;
// input order is whatever the input file holds (the caller pins the permutation)
// instead of a time-seeded shuffle; for Mash + NJ the missing allocateDeviceArrays +
// sketchConstructionOnGpu calls are inserted (reference bug, tree_generation.cu:576-587).
//
// usage: dipper_ref <mode> <input.bin> <out_prefix> [dist_type] [k] [backbone.nwk | backbone_size]
//   input.bin: int64 n, int64 bits (4 = aligned 4-bit, 2 = unaligned 2-bit),
//              uint64 len[n], then each sequence's packed words back to back.
//   modes: msa_rows   -> <out>.rows   lower-triangle distances (row i: i doubles)
//          msa_nj     -> <out>.nwk    conventional NJ tree from aligned input
//          msa_place  -> <out>.nwk    k-closest placement tree from aligned input
//          msa_place_exact / mash_place_exact -> <out>.nwk  exact placement mode (src/placement.cu)
//          mash_sketch-> <out>.sk     uint64 [n][1000] sketches
//          mash_rows  -> <out>.rows
//          mash_nj    -> <out>.nwk
//          mash_place -> <out>.nwk
//          msa_add / mash_add -> <out>.nwk   add-tips (-m 1 --add): argv[6] = backbone Newick file; input.bin holds the
//                        backbone tips first, in the leaf order of the Newick (src/tree_generation.cu:271-282 idMap),
//                        then the queries; drives initializeDeviceArrays(Tree*) + addQuery
//                        (src/placement_close_k.cu:126-264,858-990); <out>.arrays = int32 head[2n], e[8n], nxt[8n],
//                        belong[8n], double len[8n]
//          msa_dc / mash_dc  -> <out>.nwk, <out>.clusters (int32 clusterID[n], -1 for backbone tips), <out>.arrays:
//                        divide and conquer (-m 3), argv[6] = backbone size (default n/20,
//                        src/tree_generation.cu:425,545); drives findBackboneTreeDC / findClustersDC /
//                        findClusterTreeDC (src/divide_and_conquer/placement_close_k.cu:731-1535) with
//                        MSADeviceArraysDC / MashDeviceArraysDC (src/divide_and_conquer/msa.cu, mash.cu); <out>.arrays
//                        also carries int32 closest_id[20n], double closest_dis[20n] in these modes
//          msa_dc_rows -> <out>.rows  rows through MSADeviceArraysDC::distConstructionOnGpuForBackboneDC with the
//                        whole input as backbone: the well-formed twins of the six distance models
//                        (src/divide_and_conquer/msa.cu:219-264; src/MSA.cu:239-265 is broken for models 3-6)
//   every mode prints one JSON line with wall-clock phase times (cudaDeviceSynchronize
//   on both sides) to stdout.
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <string>
#include <vector>
#include <cuda_runtime.h>
#include "mash_placement.cuh"

// -DDIPPER_REF_SHIM: the same main linked against integration/mash_placement_b200.cpp + libdipper_b200.so instead of
// the reference's kernel objects (oracle/_ref/dipper_ref_b200, tests/test_shim_gpu.py); only the probe mode, which
// reaches into the reference's internals, is left out.
#ifndef DIPPER_REF_SHIM
// defined (external linkage) in src/divide_and_conquer/placement_close_k.cu:1113-1134; used by the mash_dc_probe mode only
__global__ void rearrangeHashListInClusterDC(int numSequences, int sketchSize, uint64_t* original, uint64_t* target);
#endif

#ifdef DIPPER_REF_SHIM
#define DIPPER_DRIVER_IMPL "reference-main+libdipper_b200"
#else
#define DIPPER_DRIVER_IMPL "reference-cuda"
#endif
using Clock = std::chrono::high_resolution_clock;
static double ms_since(Clock::time_point t0) {
    cudaDeviceSynchronize();
    return std::chrono::duration<double, std::milli>(Clock::now() - t0).count();
}

int main(int argc, char** argv) {
    if (argc < 4) { fprintf(stderr, "usage: dipper_ref <mode> <input.bin> <out_prefix> [dist_type] [k]\n"); return 2; }
    std::string mode = argv[1], in = argv[2], out = argv[3];
    int distType = argc > 4 ? atoi(argv[4]) : 2;
    int k = argc > 5 ? atoi(argv[5]) : 15;
    std::string extra = argc > 6 ? argv[6] : "";
    if (cudaSetDevice(0) != cudaSuccess) { fprintf(stderr, "no CUDA device\n"); return 3; }
    FILE* f = fopen(in.c_str(), "rb");
    if (!f) { fprintf(stderr, "cannot open %s\n", in.c_str()); return 2; }
    int64_t n = 0, bits = 0;
    if (fread(&n, 8, 1, f) != 1 || fread(&bits, 8, 1, f) != 1) return 2;
    std::vector<uint64_t> lens(n);
    if (fread(lens.data(), 8, n, f) != (size_t)n) return 2;
    uint64_t** seqs = new uint64_t*[n];
    int per = bits == 4 ? 16 : 32;
    for (int64_t i = 0; i < n; i++) {
        size_t w = (lens[i] + per - 1) / per;
        seqs[i] = new uint64_t[w + 1];
        seqs[i][w] = 0;
        if (fread(seqs[i], 8, w, f) != w) return 2;
    }
    fclose(f);
    std::vector<std::string> names(n);
    for (int64_t i = 0; i < n; i++) names[i] = "T" + std::to_string(i + 1);

    bool msa = mode.rfind("msa_", 0) == 0;
    MashPlacement::Param params(k, 1000, 1, distType, msa ? "m" : "r", "t");
    double t_alloc = 0, t_sketch = 0, t_dist = 0, t_tree = 0;
    cudaDeviceSynchronize();
    auto t0 = Clock::now();
    auto dump_arrays = [&](const char* suffix, int* d_head, int* d_e, int* d_nxt, int* d_belong, double* d_len, int* d_cid = nullptr,
                           double* d_cdis = nullptr) {
        std::vector<int> hh(2 * n), he(8 * n), hn(8 * n), hb(8 * n);
        std::vector<double> hl(8 * n);
        cudaMemcpy(hh.data(), d_head, sizeof(int) * 2 * n, cudaMemcpyDeviceToHost);
        cudaMemcpy(he.data(), d_e, sizeof(int) * 8 * n, cudaMemcpyDeviceToHost);
        cudaMemcpy(hn.data(), d_nxt, sizeof(int) * 8 * n, cudaMemcpyDeviceToHost);
        cudaMemcpy(hb.data(), d_belong, sizeof(int) * 8 * n, cudaMemcpyDeviceToHost);
        cudaMemcpy(hl.data(), d_len, sizeof(double) * 8 * n, cudaMemcpyDeviceToHost);
        FILE* o = fopen((out + suffix).c_str(), "wb");
        fwrite(hh.data(), 4, hh.size(), o); fwrite(he.data(), 4, he.size(), o); fwrite(hn.data(), 4, hn.size(), o);
        fwrite(hb.data(), 4, hb.size(), o); fwrite(hl.data(), 8, hl.size(), o);
        if (d_cid && d_cdis) {   // closest-leaf lists, 5 per slot for the first 4n slots (the structs allocate 20n entries)
            std::vector<int> hc(20 * n);
            std::vector<double> hd(20 * n);
            cudaMemcpy(hc.data(), d_cid, sizeof(int) * 20 * n, cudaMemcpyDeviceToHost);
            cudaMemcpy(hd.data(), d_cdis, sizeof(double) * 20 * n, cudaMemcpyDeviceToHost);
            fwrite(hc.data(), 4, hc.size(), o); fwrite(hd.data(), 8, hd.size(), o);
        }
        fclose(o);
    };
#ifndef DIPPER_REF_SHIM
    if (mode == "mash_dc_probe") {
        // Diagnostic: the stage-3 distance call of findClusterTreeDC in isolation.  argv[6] = backbone size, argv[7] =
        // probe file: int32 nb, batch[nb] (tips of one stage-3 batch, in order), int32 r (row tip = batch position),
        // int32 m, ids[m], maps[m] (leaf ids and their d_leafMap entries).  Prints d(row tip, ids[k]) as computed by
        // transferMashClusterInfoDC's steps + MashDeviceArraysDC::distSpecialIDConstructionOnGpuDC.
        int B = atoi(extra.c_str());
        params.batchSize = B; params.backboneSize = B;
        auto& md = MashPlacement::mashDeviceArraysDC;
        md.allocateDeviceArraysDC(seqs, lens.data(), n, params);
        md.sketchConstructionOnGpuDC(params, seqs, lens.data(), n);
        FILE* pf = fopen(argv[7], "rb");
        int nb = 0, r = 0, m = 0;
        fread(&nb, 4, 1, pf);
        std::vector<int> batch(nb);
        fread(batch.data(), 4, nb, pf);
        fread(&r, 4, 1, pf); fread(&m, 4, 1, pf);
        std::vector<int> ids(m), maps(m);
        fread(ids.data(), 4, m, pf); fread(maps.data(), 4, m, pf);
        fclose(pf);
        std::vector<uint64_t> local((size_t)B * 1000, 0);
        for (int l = 0; l < nb; l++) memcpy(local.data() + (size_t)l * 1000, md.h_hashList + (size_t)batch[l] * 1000, 8000);
        cudaMemcpy(md.d_hashListConst, local.data(), local.size() * 8, cudaMemcpyHostToDevice);
        uint64_t* tmp;
        cudaMalloc(&tmp, local.size() * 8);
        rearrangeHashListInClusterDC<<<1024, 1024>>>(B, 1000, md.d_hashListConst, tmp);
        std::swap(md.d_hashListConst, tmp);
        cudaFree(tmp);
        int *d_id, *d_map;
        double* d_dist;
        cudaMalloc(&d_id, 4 * m); cudaMalloc(&d_map, 4 * m); cudaMalloc(&d_dist, 8 * n);
        cudaMemset(d_dist, 0, 8 * n);
        cudaMemcpy(d_id, ids.data(), 4 * m, cudaMemcpyHostToDevice);
        cudaMemcpy(d_map, maps.data(), 4 * m, cudaMemcpyHostToDevice);
        md.distSpecialIDConstructionOnGpuDC(params, r, d_dist, m, d_id, d_map);
        std::vector<double> hd(n);
        cudaMemcpy(hd.data(), d_dist, 8 * n, cudaMemcpyDeviceToHost);
        printf("{\"probe\": [");
        for (int k = 0; k < m; k++) printf("%s%.17g", k ? ", " : "", hd[ids[k]]);
        printf("], \"cuda_status\": \"%s\"}\n", cudaGetErrorString(cudaDeviceSynchronize()));
        return 0;
    }
#endif
    if (mode == "msa_dc" || mode == "mash_dc" || mode == "msa_dc_rows") {
        // ---- divide and conquer structs (src/tree_generation.cu:422-449 aligned, :541-575 unaligned)
        const bool rows_only = mode == "msa_dc_rows";
        int B = rows_only ? (int)n : (extra.empty() ? (int)(n / 20) : atoi(extra.c_str()));
        params.batchSize = B;
        params.backboneSize = B;
        if (msa) MashPlacement::msaDeviceArraysDC.allocateDeviceArraysDC(seqs, lens.data(), n, params);
        else MashPlacement::mashDeviceArraysDC.allocateDeviceArraysDC(seqs, lens.data(), n, params);
        t_alloc = ms_since(t0);
        if (!msa) {
            t0 = Clock::now();
            MashPlacement::mashDeviceArraysDC.sketchConstructionOnGpuDC(params, seqs, lens.data(), n);
            t_sketch = ms_since(t0);
        }
        if (rows_only) {
            double* d_row;
            cudaMalloc(&d_row, sizeof(double) * n);
            std::vector<double> h(n);
            FILE* o = fopen((out + ".rows").c_str(), "wb");
            t0 = Clock::now();
            for (int i = 1; i < n; i++) {
                MashPlacement::msaDeviceArraysDC.distConstructionOnGpuForBackboneDC(params, i, d_row);
                cudaMemcpy(h.data(), d_row, sizeof(double) * i, cudaMemcpyDeviceToHost);
                fwrite(h.data(), 8, i, o);
            }
            t_dist = ms_since(t0);
            fclose(o);
        } else {
            std::ofstream os(out + ".nwk");
            auto& kp = MashPlacement::kplacementDeviceArraysDC;
            kp.allocateDeviceArraysDC(B, n);
            t0 = Clock::now();
            kp.findBackboneTreeDC(params, MashPlacement::mashDeviceArraysDC, MashPlacement::matrixReader,
                                  MashPlacement::msaDeviceArraysDC, MashPlacement::kplacementDeviceArraysHostDC);
            t_dist = ms_since(t0);   // "dist_ms" = backbone stage in this mode
            t0 = Clock::now();
            kp.findClustersDC(params, MashPlacement::mashDeviceArraysDC, MashPlacement::matrixReader,
                              MashPlacement::msaDeviceArraysDC, MashPlacement::kplacementDeviceArraysHostDC);
            t_sketch += ms_since(t0);   // "sketch_ms" (+) = cluster assignment stage in this mode
            {
                std::vector<int> cl(n, -1);
                for (int64_t j = B; j < n; j++) cl[j] = MashPlacement::kplacementDeviceArraysHostDC.clusterID[j];
                FILE* o = fopen((out + ".clusters").c_str(), "wb");
                fwrite(cl.data(), 4, cl.size(), o);
                fclose(o);
            }
            t0 = Clock::now();
            kp.findClusterTreeDC(params, MashPlacement::mashDeviceArraysDC, MashPlacement::matrixReader,
                                 MashPlacement::msaDeviceArraysDC, MashPlacement::kplacementDeviceArraysHostDC);
            t_tree = ms_since(t0);
            kp.printTreeDC(names, os);
            dump_arrays(".arrays", kp.d_head, kp.d_e, kp.d_nxt, kp.d_belong, kp.d_len, kp.d_closest_id, kp.d_closest_dis);
        }
        cudaError_t e = cudaDeviceSynchronize();
        printf("{\"impl\": \"" DIPPER_DRIVER_IMPL "\", \"mode\": \"%s\", \"n\": %lld, \"backbone\": %d, \"alloc_ms\": %.3f, \"assign_ms\": %.3f, "
               "\"backbone_ms\": %.3f, \"clusters_ms\": %.3f, \"cuda_status\": \"%s\"}\n",
               mode.c_str(), (long long)n, B, t_alloc, t_sketch, t_dist, t_tree, cudaGetErrorString(e));
        return e == cudaSuccess ? 0 : 4;
    }
    if (msa) MashPlacement::msaDeviceArrays.allocateDeviceArrays(seqs, lens.data(), n, params);
    else MashPlacement::mashDeviceArrays.allocateDeviceArrays(seqs, lens.data(), n, params);
    t_alloc = ms_since(t0);
    if (!msa) {
        t0 = Clock::now();
        MashPlacement::mashDeviceArrays.sketchConstructionOnGpu(params);
        t_sketch = ms_since(t0);
    }
    if (mode == "mash_sketch") {
        // h_hashList holds the transposed layout hash[t*n + seq] (src/mash.cu:381-383,418-419)
        std::vector<uint64_t> sk((size_t)n * 1000);
        for (int64_t s = 0; s < n; s++)
            for (int t = 0; t < 1000; t++) sk[s * 1000 + t] = MashPlacement::mashDeviceArrays.h_hashList[(size_t)t * n + s];
        FILE* o = fopen((out + ".sk").c_str(), "wb");
        fwrite(sk.data(), 8, sk.size(), o);
        fclose(o);
    } else if (mode == "msa_rows" || mode == "mash_rows") {
        double* d_row;
        cudaMalloc(&d_row, sizeof(double) * n);
        std::vector<double> h(n);
        FILE* o = fopen((out + ".rows").c_str(), "wb");
        t0 = Clock::now();
        for (int i = 1; i < n; i++) {
            if (msa) MashPlacement::msaDeviceArrays.distConstructionOnGpu(params, i, d_row);
            else MashPlacement::mashDeviceArrays.distConstructionOnGpu(params, i, d_row);
            cudaMemcpy(h.data(), d_row, sizeof(double) * i, cudaMemcpyDeviceToHost);
            fwrite(h.data(), 8, i, o);
        }
        t_dist = ms_since(t0);
        fclose(o);
    } else if (mode == "msa_nj" || mode == "mash_nj") {
        std::ofstream os(out + ".nwk");
        t0 = Clock::now();
        MashPlacement::njDeviceArrays.getDismatrix((int)n, params, MashPlacement::mashDeviceArrays,
                                                   MashPlacement::matrixReader, MashPlacement::msaDeviceArrays);
        t_dist = ms_since(t0);
        t0 = Clock::now();
        MashPlacement::njDeviceArrays.findNeighbourJoiningTree(names, os);
        t_tree = ms_since(t0);
    } else if (mode == "msa_place" || mode == "mash_place") {
        std::ofstream os(out + ".nwk");
        MashPlacement::kplacementDeviceArrays.allocateDeviceArrays(n);
        t0 = Clock::now();
        MashPlacement::kplacementDeviceArrays.findPlacementTree(params, MashPlacement::mashDeviceArrays,
                                                                MashPlacement::matrixReader, MashPlacement::msaDeviceArrays);
        t_tree = ms_since(t0);
        MashPlacement::kplacementDeviceArrays.printTree(names, os);
    } else if (mode == "msa_add" || mode == "mash_add") {
        // src/tree_generation.cu:252-332: Tree(newick, #seqs), leaves numbered by order of appearance; the caller
        // already ordered input.bin that way (idMap), queries follow the backbone tips
        std::ifstream tf(extra);
        if (!tf) { fprintf(stderr, "cannot open backbone tree %s\n", extra.c_str()); return 2; }
        std::string nwk;
        std::getline(tf, nwk);
        Tree* t = new Tree(nwk, (size_t)n);
        const int B = (int)t->m_numLeaves;
        for (auto& kv : t->allNodes)
            if (kv.second->children.empty() && kv.second->idx < B) names[kv.second->idx] = kv.first;
        for (int64_t i = B; i < n; i++) names[i] = "Q" + std::to_string(i - B + 1);
        std::ofstream os(out + ".nwk");
        auto& kp = MashPlacement::kplacementDeviceArrays;
        kp.allocateDeviceArrays(n, B);
        t0 = Clock::now();
        kp.initializeDeviceArrays(t);
        t_dist = ms_since(t0);      // "dist_ms" = backbone load in this mode
        t0 = Clock::now();
        kp.addQuery(params, MashPlacement::mashDeviceArrays, MashPlacement::matrixReader, MashPlacement::msaDeviceArrays);
        t_tree = ms_since(t0);
        kp.printTree(names, os);
        dump_arrays(".arrays", kp.d_head, kp.d_e, kp.d_nxt, kp.d_belong, kp.d_len);
    } else if (mode == "msa_place_exact" || mode == "mash_place_exact") {
        std::ofstream os(out + ".nwk");
        MashPlacement::placementDeviceArrays.allocateDeviceArrays(n);
        t0 = Clock::now();
        MashPlacement::placementDeviceArrays.findPlacementTree(params, MashPlacement::mashDeviceArrays,
                                                               MashPlacement::matrixReader, MashPlacement::msaDeviceArrays);
        t_tree = ms_since(t0);
        MashPlacement::placementDeviceArrays.printTree(names, os);
    } else {
        fprintf(stderr, "unknown mode %s\n", mode.c_str());
        return 2;
    }
    cudaError_t e = cudaDeviceSynchronize();
    printf("{\"impl\": \"" DIPPER_DRIVER_IMPL "\", \"mode\": \"%s\", \"n\": %lld, \"alloc_ms\": %.3f, \"sketch_ms\": %.3f, "
           "\"dist_ms\": %.3f, \"tree_ms\": %.3f, \"cuda_status\": \"%s\"}\n",
           mode.c_str(), (long long)n, t_alloc, t_sketch, t_dist, t_tree, cudaGetErrorString(e));
    return e == cudaSuccess ? 0 : 4;
}
