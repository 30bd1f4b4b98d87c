// Several GPUs of one box driven by ONE process: one dipb_ctx and one host thread per device, the
// shards exchanged with peer copies over NVLink (cudaMemcpy2DAsync between devices) or, where they are a few bytes per
// tip, through host memory.  The reference is single GPU (device 1 hard-coded, src/tree_generation.cu:240).
//   * distance matrices: 128-row-aligned row blocks balanced by triangle area; device d computes rows [r0, r1) x
//     columns [0, r1), device 0 pulls exactly that trapezoid from every peer and mirrors it;
//   * divide and conquer (-m 3): the backbone placement is deterministic and runs on every device, the queries of
//     stage 2 (the bulk of the work) are split evenly, the cluster ids (one int per tip) meet on the host, device 0
//     places the clusters.
// bench.py / tools/dc_multi_gpu.py keep the one-process-per-GPU form over torch.distributed that the driver launches.
#include <algorithm>
#include <chrono>
#include <cmath>
#include <functional>
#include <string>
#include <thread>
#include <vector>
#include "common.cuh"
#include "msa.cuh"

struct dipb_dc_state;
extern "C" {
int dipb_dc_begin(dipb_ctx* ctx, const dipb_dist_source* src, int n, int backbone, dipb_dc_state** out);
int dipb_dc_assign(dipb_dc_state* st, int q0, int q1, int32_t* h_cluster);
int dipb_dc_set_clusters(dipb_dc_state* st, const int32_t* h_cluster_all, int* num_clusters);
int dipb_dc_run_clusters(dipb_dc_state* st, int c0, int c1);
int dipb_dc_finish(dipb_dc_state* st, dipb_tree** out);
}

struct dipb_multi {
    std::vector<dipb_ctx*> ctx;
    std::vector<dipb_msa*> msa;
    // row blocks of the peers for the matrix gather: plain cudaMalloc memory (peer readable once peer access is on;
    // memory of the stream-ordered pools is private to its device, and a pool opened to a peer with cudaMemPoolSetAccess
    // refused to grow while that peer was busy), kept across calls
    std::vector<double*> block;
    std::vector<size_t> block_bytes;
    std::vector<int32_t> clusters;      // cluster ids of the last dipb_multi_dc (test hook)
    double t_ms[4] = {0, 0, 0, 0};      // last call: compute (max over devices), gather, mirror, total
};

using namespace dipb;

// Device 0 pulls rows [r0, r1) of a peer's matrix (entries below the diagonal only) straight out of the peer's memory
// over NVLink and writes them twice: in place and mirrored above the diagonal.  32 x 32 tiles through shared memory so
// that both the peer reads and the two local writes are coalesced.  (A pitched cudaMemcpy3DPeerAsync of the same
// trapezoid ran at 36 GB/s; this kernel is bound by the NVLink read.)
__global__ void gather_mirror_kernel(const double* __restrict__ src, size_t lds, double* __restrict__ dst, int n, int r0, int r1) {
    __shared__ double tile[32][33];
    const int i0 = r0 + blockIdx.y * 32, j0 = blockIdx.x * 32;
    if (j0 > i0 + 31) return;                       // tile entirely above the diagonal
    const int tx = threadIdx.x, ty = threadIdx.y;   // 32 x 8
    for (int k = ty; k < 32; k += 8) {
        const int i = i0 + k, j = j0 + tx;
        double v = 0.0;
        if (i < r1 && j < i) { v = src[(size_t)(i - r0) * lds + j]; dst[(size_t)i * n + j] = v; }   // src row 0 = matrix row r0
        tile[k][tx] = v;
    }
    __syncthreads();
    for (int k = ty; k < 32; k += 8) {
        const int j = j0 + k, i = i0 + tx;          // write dst[j][i], coalesced along i
        if (i < r1 && j < i) dst[(size_t)j * n + i] = tile[tx][k];
    }
}

namespace {
// run fn(d) on one host thread per device; the first failure (lowest device) wins and its message is re-raised here
int on_all(dipb_multi* m, const std::function<int(int)>& fn) {
    const int nd = (int)m->ctx.size();
    std::vector<int> rc(nd, 0);
    std::vector<std::string> msg(nd);
    std::vector<std::thread> th;
    for (int d = 0; d < nd; d++)
        th.emplace_back([&, d]() {
            cudaSetDevice(m->ctx[d]->device);
            rc[d] = fn(d);
            if (rc[d]) msg[d] = dipb_last_error();
        });
    for (auto& t : th) t.join();
    for (int d = 0; d < nd; d++)
        if (rc[d]) { set_error("device %d: %s", m->ctx[d]->device, msg[d].c_str()); return rc[d]; }
    return 0;
}
// rows [r0, r1) of device d: 128-aligned cuts at equal triangle area (rows near the bottom are longer)
void row_shard(int n, int nd, int d, int* r0, int* r1) {
    auto cut = [&](int k) {
        if (k <= 0) return 0;
        if (k >= nd) return n;
        int r = (int)std::llround(std::sqrt((double)k / nd) * n / 128.0) * 128;
        return std::min(std::max(r, 0), n);
    };
    *r0 = cut(d); *r1 = cut(d + 1);
}
}  // namespace

extern "C" {

int dipb_multi_init(const int* devices, int n_devices, dipb_multi** out) {
    if (!devices || n_devices < 1 || !out) { set_error("dipb_multi_init: bad argument"); return DIPB_E_ARG; }
    dipb_multi* m = new dipb_multi();
    for (int d = 0; d < n_devices; d++) {
        dipb_ctx* c = nullptr;
        int rc = dipb_init(devices[d], &c);
        if (rc) { for (auto* x : m->ctx) dipb_destroy(x); delete m; return rc; }
        m->ctx.push_back(c);
    }
    // device 0 pulls from every peer
    for (int d = 1; d < n_devices; d++) {
        int can = 0;
        cudaDeviceCanAccessPeer(&can, devices[0], devices[d]);
        if (can) {
            cudaSetDevice(devices[0]);
            cudaError_t e = cudaDeviceEnablePeerAccess(devices[d], 0);
            if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) { set_error("dipb_multi_init: peer access %d -> %d: %s", devices[0], devices[d], cudaGetErrorString(e)); for (auto* x : m->ctx) dipb_destroy(x); delete m; return DIPB_E_CUDA; }
            cudaGetLastError();
        } else { set_error("dipb_multi_init: device %d cannot access device %d (no NVLink / PCIe peer path)", devices[0], devices[d]); for (auto* x : m->ctx) dipb_destroy(x); delete m; return DIPB_E_UNSUPPORTED; }
    }
    m->msa.assign(n_devices, nullptr);
    m->block.assign(n_devices, nullptr);
    m->block_bytes.assign(n_devices, 0);
    *out = m;
    return 0;
}

void dipb_multi_destroy(dipb_multi* m) {
    if (!m) return;
    for (size_t d = 0; d < m->ctx.size(); d++) {
        if (m->msa[d]) dipb_msa_free(m->msa[d]);
        if (m->block[d]) { cudaSetDevice(m->ctx[d]->device); cudaFree(m->block[d]); }
        dipb_destroy(m->ctx[d]);
    }
    delete m;
}

int dipb_multi_devices(const dipb_multi* m) { return m ? (int)m->ctx.size() : 0; }
dipb_ctx* dipb_multi_ctx(dipb_multi* m, int d) { return (m && d >= 0 && d < (int)m->ctx.size()) ? m->ctx[d] : nullptr; }
double dipb_multi_elapsed_ms(const dipb_multi* m, int what) { return (m && what >= 0 && what < 4) ? m->t_ms[what] : -1.0; }

// every device gets the packed sequences (MSADeviceArrays::allocateDeviceArrays on each, src/MSA.cu:14-72)
int dipb_multi_msa_upload_flat(dipb_multi* m, const uint64_t* flat, size_t n, uint64_t seq_len) {
    if (!m || !flat) { set_error("dipb_multi_msa_upload_flat: bad argument"); return DIPB_E_ARG; }
    for (auto*& x : m->msa) { if (x) dipb_msa_free(x); x = nullptr; }
    return on_all(m, [&](int d) { return dipb_msa_upload_flat(m->ctx[d], flat, n, seq_len, &m->msa[d]); });
}

// NJDeviceArrays::getDismatrix (src/neighborJoining.cu:35-85) over several devices: the full mirrored matrix on device 0
int dipb_multi_msa_dist_matrix(dipb_multi* m, int dist_type, dipb_matrix** out) {
    if (!m || !out || !m->msa[0]) { set_error("dipb_multi_msa_dist_matrix: upload the sequences first"); return DIPB_E_STATE; }
    const int nd = (int)m->ctx.size(), n = m->msa[0]->n;
    dipb_matrix* M0 = nullptr;
    std::vector<size_t> lds(nd, 0);
    auto t0 = std::chrono::steady_clock::now();
    int rc = on_all(m, [&](int d) {
        int r0, r1;
        row_shard(n, nd, d, &r0, &r1);
        int r = 0;
        if (d == 0) r = dipb_msa_dist_matrix_rows(m->msa[0], dist_type, 0, nd > 1 ? r1 : n, &M0);   // full matrix, own rows mirrored
        else if (r1 > r0) {
            // rows [r0, r1) x columns [0, r1) as a rectangular block in peer-readable memory
            lds[d] = (size_t)((r1 + 127) / 128 * 128);
            const size_t need = (size_t)(r1 - r0) * lds[d] * sizeof(double);
            if (need > m->block_bytes[d]) {
                if (m->block[d]) cudaFree(m->block[d]);
                m->block[d] = nullptr; m->block_bytes[d] = 0;
                if (cudaMalloc(&m->block[d], need) != cudaSuccess) { set_error("dipb_multi_msa_dist_matrix: %zu MB row block: %s", need >> 20, cudaGetErrorString(cudaGetLastError())); return (int)DIPB_E_NOMEM; }
                m->block_bytes[d] = need;
            }
            r = dipb_msa_dist_block(m->msa[d], dist_type, r0, r1, r1, m->block[d], lds[d]);
        }
        if (!r) r = dipb_sync(m->ctx[d]);
        return r;
    });
    auto t1 = std::chrono::steady_clock::now();
    if (!rc) {
        // device 0 pulls the rows of every peer below the diagonal and mirrors them on the way in
        dipb_ctx* c0 = m->ctx[0];
        cudaSetDevice(c0->device);
        for (int d = 1; d < nd && !rc; d++) {
            int r0, r1;
            row_shard(n, nd, d, &r0, &r1);
            if (r1 <= r0) continue;
            dim3 grid((r1 + 31) / 32, (r1 - r0 + 31) / 32), block(32, 8);
            gather_mirror_kernel<<<grid, block, 0, c0->stream>>>(m->block[d], lds[d], M0->d, n, r0, r1);
            c0->launches++;
            if (cudaGetLastError() != cudaSuccess) { set_error("dipb_multi_msa_dist_matrix: gather from device %d failed to launch", m->ctx[d]->device); rc = DIPB_E_CUDA; }
        }
        if (!rc && cudaStreamSynchronize(c0->stream) != cudaSuccess) { set_error("dipb_multi_msa_dist_matrix: gather failed: %s", cudaGetErrorString(cudaGetLastError())); rc = DIPB_E_CUDA; }
    }
    auto t2 = std::chrono::steady_clock::now();
    auto t3 = t2;
    cudaSetDevice(m->ctx[0]->device);
    if (rc) { if (M0) dipb_matrix_free(M0); return rc; }
    auto ms = [](auto a, auto b) { return std::chrono::duration<double, std::milli>(b - a).count(); };
    m->t_ms[0] = ms(t0, t1); m->t_ms[1] = ms(t1, t2); m->t_ms[2] = ms(t2, t3); m->t_ms[3] = ms(t0, t3);
    *out = M0;
    return 0;
}

// -m 3 over several devices (findBackboneTreeDC / findClustersDC / findClusterTreeDC,
// src/divide_and_conquer/placement_close_k.cu:731-1535): stage 2 sharded by query range, tree on device 0
int dipb_multi_dc(dipb_multi* m, int dist_type, int backbone, dipb_tree** out) {
    if (!m || !out || !m->msa[0]) { set_error("dipb_multi_dc: upload the sequences first"); return DIPB_E_STATE; }
    const int nd = (int)m->ctx.size(), n = m->msa[0]->n;
    if (backbone < 2 || backbone >= n) { set_error("dipb_multi_dc: backbone size %d must be in [2, n)", backbone); return DIPB_E_ARG; }
    std::vector<dipb_dc_state*> st(nd, nullptr);
    std::vector<int32_t> cl((size_t)n, -1);
    const int nq = n - backbone;
    auto t0 = std::chrono::steady_clock::now();
    int rc = on_all(m, [&](int d) {
        dipb_dist_source src{};
        src.msa = m->msa[d]; src.dist_type = dist_type;
        int r = dipb_dc_begin(m->ctx[d], &src, n, backbone, &st[d]);          // stage 1, identical on every device
        if (r) return r;
        const int q0 = backbone + (int)((long long)nq * d / nd), q1 = backbone + (int)((long long)nq * (d + 1) / nd);
        return dipb_dc_assign(st[d], q0, q1, cl.data() + q0);                 // stage 2, this device's queries
    });
    auto t1 = std::chrono::steady_clock::now();
    for (int d = 1; d < nd; d++) if (st[d]) { cudaSetDevice(m->ctx[d]->device); dipb_dc_finish(st[d], nullptr); st[d] = nullptr; }
    cudaSetDevice(m->ctx[0]->device);
    int nc = 0;
    if (!rc) rc = dipb_dc_set_clusters(st[0], cl.data(), &nc);
    if (!rc) rc = dipb_dc_run_clusters(st[0], 0, nc);                         // stage 3 on device 0
    if (rc) { if (st[0]) dipb_dc_finish(st[0], nullptr); return rc; }
    rc = dipb_dc_finish(st[0], out);
    auto t2 = std::chrono::steady_clock::now();
    auto ms = [](auto a, auto b) { return std::chrono::duration<double, std::milli>(b - a).count(); };
    m->t_ms[0] = ms(t0, t1); m->t_ms[1] = 0; m->t_ms[2] = ms(t1, t2); m->t_ms[3] = ms(t0, t2);
    m->clusters = cl;
    return rc;
}

int dipb_multi_dc_cluster_ids(const dipb_multi* m, int32_t* h_out, int n) {
    if (!m || !h_out || (int)m->clusters.size() != n) { set_error("dipb_multi_dc_cluster_ids: no matching dipb_multi_dc run"); return DIPB_E_STATE; }
    memcpy(h_out, m->clusters.data(), sizeof(int32_t) * (size_t)n);
    return 0;
}

}  // extern "C"
