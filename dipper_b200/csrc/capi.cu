// C ABI glue of libdipper_b200 (include/dipper_b200.h): contexts, uploads, matrices.
#include <vector>
#include "common.cuh"
#include "msa.cuh"
#include "nj.cuh"

namespace dipb {
static thread_local char g_err[1024] = "";
void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof g_err, fmt, ap);
    va_end(ap);
}
void ctx_release(dipb_ctx* c) {
    if (!c || c->refs.fetch_sub(1, std::memory_order_acq_rel) != 1) return;
    cudaSetDevice(c->device);
    if (c->stream) cudaStreamSynchronize(c->stream);   // frees queued by the children are stream ordered
    if (c->stage) cudaFree(c->stage);
    if (c->ev0) cudaEventDestroy(c->ev0);
    if (c->ev1) cudaEventDestroy(c->ev1);
    if (c->stream) cudaStreamDestroy(c->stream);
    delete c;
}
}  // namespace dipb

using namespace dipb;

extern "C" {

const char* dipb_last_error(void) { return g_err; }
const char* dipb_version(void) { return "dipper_b200 0.1 (sm_100a)"; }

int dipb_init(int device, dipb_ctx** out) {
    if (!out) { set_error("dipb_init: null out"); return DIPB_E_ARG; }
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0) {
        set_error("dipb_init: no CUDA device (%s); this library has no CPU fallback",
                  e != cudaSuccess ? cudaGetErrorString(e) : "device count 0");
        return DIPB_E_CUDA;
    }
    if (device < 0 || device >= count) { set_error("dipb_init: device %d out of range (%d visible)", device, count); return DIPB_E_ARG; }
    DIPB_CUDA(cudaSetDevice(device));
    dipb_ctx* c = new dipb_ctx();
    c->device = device;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) { set_error("dipb_init: cudaGetDeviceProperties failed"); delete c; return DIPB_E_CUDA; }
    c->num_sms = prop.multiProcessorCount;
    if (prop.major < 10) {
        set_error("dipb_init: device %d is sm_%d%d; this library is built for sm_100a only", device, prop.major, prop.minor);
        delete c;
        return DIPB_E_CUDA;
    }
    if (cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking) != cudaSuccess || cudaEventCreate(&c->ev0) != cudaSuccess ||
        cudaEventCreate(&c->ev1) != cudaSuccess) {
        set_error("dipb_init: stream / event creation failed: %s", cudaGetErrorString(cudaGetLastError()));
        ctx_release(c);
        return DIPB_E_CUDA;
    }
    {
        // keep freed pool pages for re-use (see pool_alloc in common.cuh)
        cudaMemPool_t pool;
        if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) {
            unsigned long long keep = ~0ull;
            cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
        }
        cudaGetLastError();
    }
    for (int i = 0; i < DIPB_T_COUNT; i++) c->elapsed[i] = -1.0;
    *out = c;
    return DIPB_OK;
}

void dipb_destroy(dipb_ctx* c) {
    if (!c || c->destroyed) return;
    c->destroyed = true;
    ctx_release(c);
}
int dipb_ctx_refs(const dipb_ctx* c) { return c ? c->refs.load() : 0; }

double dipb_elapsed_ms(dipb_ctx* c, int what) {
    if (!c || what < 0 || what >= DIPB_T_COUNT) return -1.0;
    return c->elapsed[what];
}
uint64_t dipb_kernel_launches(dipb_ctx* c) { return c ? c->launches : 0; }
int dipb_sync(dipb_ctx* c) {
    if (!c) return DIPB_E_ARG;
    DIPB_CUDA(cudaStreamSynchronize(c->stream));
    return 0;
}

// ---- aligned MSA -------------------------------------------------------------
static int msa_create_impl(dipb_msa* m, const uint64_t* d_in, size_t n, uint64_t seq_len) {
    dipb_ctx* c = m->ctx;
    m->n = (int)n;
    m->seq_len = (int)seq_len;
    m->npad = (int)((n + MSA_TS - 1) / MSA_TS * MSA_TS);
    int w32 = (int)((seq_len + 31) / 32);
    m->nkc = (w32 + MSA_KC - 1) / MSA_KC;
    if (m->nkc < 1) m->nkc = 1;
    size_t words = (size_t)(m->npad / MSA_TS) * m->nkc * MSA_SLAB_WORDS;
    DIPB_CUDA(pool_alloc(c, (void**)&m->planes, words * sizeof(uint32_t)));
    DIPB_CUDA(pool_alloc(c, (void**)&m->nv, sizeof(int) * m->npad));
    DIPB_CUDA(cudaMemsetAsync(m->nv, 0, sizeof(int) * m->npad, c->stream));
    return msa_repack(m, d_in, (int)((seq_len + 15) / 16));
}
static int msa_create(dipb_ctx* c, const uint64_t* d_in, size_t n, uint64_t seq_len, dipb_msa** out) {
    dipb_msa* m = new dipb_msa();
    m->ctx = c;
    ctx_retain(c);
    const int rc = msa_create_impl(m, d_in, n, seq_len);
    if (rc) { dipb_msa_free(m); return rc; }   // nothing leaks on a failed upload
    *out = m;
    return 0;
}

int dipb_msa_upload_flat(dipb_ctx* c, const uint64_t* flat, size_t n, uint64_t seq_len, dipb_msa** out) {
    if (!c || !flat || !out || n == 0 || seq_len == 0) { set_error("dipb_msa_upload_flat: bad argument"); return DIPB_E_ARG; }
    if (seq_len > 0x7fffffffULL || n > 0x7fffff00ULL) { set_error("dipb_msa_upload_flat: too large"); return DIPB_E_ARG; }
    DIPB_CUDA(cudaSetDevice(c->device));
    size_t comp = (seq_len + 15) / 16;
    uint64_t* d_in = nullptr;
    // Staging block.  Deliberately NOT from the stream-ordered pool: a short-lived 450 MB block there splits the free 7 GB
    // block the next matrix wants and makes the pool grow again (measured 90-270 ms per tree).  Up to 1 GB it is kept with
    // the context and reused by the next upload; larger inputs allocate and free their own.
    const size_t in_bytes = n * comp * sizeof(uint64_t);
    const bool cached = in_bytes <= ((size_t)1 << 30);
    if (cached) {
        if (c->stage_bytes < in_bytes) {
            if (c->stage) { cudaStreamSynchronize(c->stream); cudaFree(c->stage); c->stage = nullptr; c->stage_bytes = 0; }
            DIPB_CUDA(cudaMalloc(&c->stage, in_bytes));
            c->stage_bytes = in_bytes;
        }
        d_in = static_cast<uint64_t*>(c->stage);
    } else {
        DIPB_CUDA(cudaMalloc(&d_in, in_bytes));
    }
    int rc = timer_begin(c);
    if (!rc && cudaMemcpyAsync(d_in, flat, n * comp * sizeof(uint64_t), cudaMemcpyHostToDevice, c->stream) != cudaSuccess) {
        set_error("dipb_msa_upload_flat: H2D copy failed: %s", cudaGetErrorString(cudaGetLastError()));
        rc = DIPB_E_CUDA;
    }
    dipb_msa* m = nullptr;
    if (!rc) rc = msa_create(c, d_in, n, seq_len, &m);
    if (!rc) rc = timer_end(c, DIPB_T_MSA_UPLOAD);
    cudaStreamSynchronize(c->stream);
    if (!cached) cudaFree(d_in);
    if (rc) { dipb_msa_free(m); return rc; }
    *out = m;
    return 0;
}

int dipb_msa_upload(dipb_ctx* c, const uint64_t* const* seq4, const uint64_t* len, size_t n, dipb_msa** out) {
    if (!c || !seq4 || !len || !out || n == 0) { set_error("dipb_msa_upload: bad argument"); return DIPB_E_ARG; }
    // seqLen = len[0], as MSADeviceArrays::allocateDeviceArrays (src/MSA.cu:19); rows are
    // flattened with their own ceil(len/16) in the reference, which only works when all
    // lengths agree, so that is required here.
    uint64_t L = len[0];
    for (size_t i = 0; i < n; i++)
        if (len[i] != L) { set_error("dipb_msa_upload: sequence %zu has length %llu, expected %llu (aligned input)", i, (unsigned long long)len[i], (unsigned long long)L); return DIPB_E_ARG; }
    size_t comp = (L + 15) / 16;
    std::vector<uint64_t> flat(n * comp);
    for (size_t i = 0; i < n; i++) memcpy(flat.data() + i * comp, seq4[i], comp * sizeof(uint64_t));
    return dipb_msa_upload_flat(c, flat.data(), n, L, out);
}

void dipb_msa_free(dipb_msa* m) {
    if (!m) return;
    cudaSetDevice(m->ctx->device);
    pool_free(m->ctx, m->planes);
    pool_free(m->ctx, m->nv);
    pool_free(m->ctx, m->tc_S);
    pool_free(m->ctx, m->tc_V);
    pool_free(m->ctx, m->tc_Sx);
    pool_free(m->ctx, m->tc_Vx);
    ctx_release(m->ctx);
    delete m;
}

int dipb_msa_drop_operands(dipb_msa* m) {
    if (!m) { set_error("dipb_msa_drop_operands: null msa"); return DIPB_E_ARG; }
    DIPB_CUDA(cudaSetDevice(m->ctx->device));
    pool_free(m->ctx, m->tc_S); pool_free(m->ctx, m->tc_V); pool_free(m->ctx, m->tc_Sx); pool_free(m->ctx, m->tc_Vx);
    m->tc_S = m->tc_V = m->tc_Sx = m->tc_Vx = nullptr;
    m->tc_rows = m->tc_have = m->tc_xrows = 0;
    return 0;
}

int dipb_msa_dist_row(dipb_msa* m, int dist_type, int row, double* d_out) {
    if (!m || !d_out) { set_error("dipb_msa_dist_row: bad argument"); return DIPB_E_ARG; }
    if (row < 0 || row >= m->n) { set_error("dipb_msa_dist_row: row %d out of range", row); return DIPB_E_ARG; }
    DIPB_CUDA(cudaSetDevice(m->ctx->device));
    if (row == 0) return 0;
    int rc = msa_block(m, dist_type, row, row + 1, row, d_out, (size_t)m->n);
    if (rc) return rc;
    DIPB_CUDA(cudaStreamSynchronize(m->ctx->stream));  // distConstructionOnGpu callers sync right after (src/placement_close_k.cu:781)
    return 0;
}

int dipb_msa_dist_row_host(dipb_msa* m, int dist_type, int row, double* h_out) {
    if (!m || !h_out) { set_error("dipb_msa_dist_row_host: bad argument"); return DIPB_E_ARG; }
    if (row < 0 || row >= m->n) { set_error("dipb_msa_dist_row_host: row %d out of range", row); return DIPB_E_ARG; }
    if (row == 0) return 0;
    DIPB_CUDA(cudaSetDevice(m->ctx->device));
    double* d = nullptr;
    DIPB_CUDA(cudaMalloc(&d, sizeof(double) * m->n));
    int rc = dipb_msa_dist_row(m, dist_type, row, d);
    if (!rc) {
        cudaError_t e = cudaMemcpy(h_out, d, sizeof(double) * row, cudaMemcpyDeviceToHost);
        if (e != cudaSuccess) { set_error("dipb_msa_dist_row_host: D2H failed: %s", cudaGetErrorString(e)); rc = DIPB_E_CUDA; }
    }
    cudaFree(d);
    return rc;
}

int dipb_msa_dist_block(dipb_msa* m, int dist_type, int r0, int r1, int ncols, double* d_out, size_t ld) {
    if (!m || !d_out) { set_error("dipb_msa_dist_block: bad argument"); return DIPB_E_ARG; }
    DIPB_CUDA(cudaSetDevice(m->ctx->device));
    int rc = timer_begin(m->ctx);
    if (rc) return rc;
    rc = msa_block(m, dist_type, r0, r1, ncols, d_out, ld);
    if (rc) return rc;
    return timer_end(m->ctx, DIPB_T_MSA_DIST);
}

int dipb_msa_counts(dipb_msa* m, int i0, int i1, int j0, int j1, int32_t* h_match, int32_t* h_useful) {
    if (!m || !h_match || !h_useful || i0 < 0 || i1 > m->n || i0 >= i1 || j0 < 0 || j1 > m->n || j0 >= j1) {
        set_error("dipb_msa_counts: bad argument");
        return DIPB_E_ARG;
    }
    DIPB_CUDA(cudaSetDevice(m->ctx->device));
    size_t ld = (size_t)(j1 + 127) / 128 * 128;
    int rows = i1 - i0;
    int *dm = nullptr, *db = nullptr;
    DIPB_CUDA(cudaMalloc(&dm, sizeof(int) * rows * ld));
    DIPB_CUDA(cudaMalloc(&db, sizeof(int) * rows * ld));
    int rc = msa_counts_dev(m, i0, i1, j1, dm, db, ld);
    if (!rc) {
        std::vector<int> hm(rows * ld), hb(rows * ld), nv(m->n);
        cudaStreamSynchronize(m->ctx->stream);
        cudaError_t e1 = cudaMemcpy(hm.data(), dm, sizeof(int) * rows * ld, cudaMemcpyDeviceToHost);
        cudaError_t e2 = cudaMemcpy(hb.data(), db, sizeof(int) * rows * ld, cudaMemcpyDeviceToHost);
        cudaError_t e3 = cudaMemcpy(nv.data(), m->nv, sizeof(int) * m->n, cudaMemcpyDeviceToHost);
        if (e1 != cudaSuccess || e2 != cudaSuccess || e3 != cudaSuccess) { set_error("dipb_msa_counts: D2H failed"); rc = DIPB_E_CUDA; }
        else
            for (int i = i0; i < i1; i++)
                for (int j = j0; j < j1; j++) {
                    size_t o = (size_t)(i - i0) * ld + j, q = (size_t)(i - i0) * (j1 - j0) + (j - j0);
                    h_match[q] = hm[o];
                    h_useful[q] = nv[i] + nv[j] - hb[o];
                }
    }
    cudaFree(dm);
    cudaFree(db);
    return rc;
}

int dipb_msa_dist_matrix_rows(dipb_msa* m, int dist_type, int row_begin, int row_end, dipb_matrix** out) {
    if (!m || !out) { set_error("dipb_msa_dist_matrix: bad argument"); return DIPB_E_ARG; }
    if (dist_type < 1 || dist_type > 6) { set_error("dipb_msa_dist_matrix: distance type %d not in 1..6", dist_type); return DIPB_E_ARG; }
    DIPB_CUDA(cudaSetDevice(m->ctx->device));
    dipb_matrix* M = new dipb_matrix();
    M->ctx = m->ctx;
    ctx_retain(M->ctx);
    M->n = m->n;
    size_t bytes = (size_t)m->n * m->n * sizeof(double);
    cudaError_t e = pool_alloc(m->ctx, (void**)&M->d, bytes);
    if (e != cudaSuccess) { set_error("dipb_msa_dist_matrix: allocation of %zu bytes failed: %s", bytes, cudaGetErrorString(e)); M->d = nullptr; dipb_matrix_free(M); return DIPB_E_NOMEM; }
    int rc = 0;
    if ((row_begin != 0 || row_end != m->n) && cudaMemsetAsync(M->d, 0, bytes, m->ctx->stream) != cudaSuccess) { set_error("dipb_msa_dist_matrix: memset failed"); rc = DIPB_E_CUDA; }
    if (!rc) rc = timer_begin(m->ctx);
    if (!rc) rc = msa_matrix(m, dist_type, row_begin, row_end, M->d);
    if (!rc) rc = timer_end(m->ctx, DIPB_T_MSA_DIST);
    if (rc) { dipb_matrix_free(M); return rc; }
    *out = M;
    return 0;
}

int dipb_msa_dist_matrix(dipb_msa* m, int dist_type, dipb_matrix** out) {
    if (!m) { set_error("dipb_msa_dist_matrix: null msa"); return DIPB_E_ARG; }
    return dipb_msa_dist_matrix_rows(m, dist_type, 0, m->n, out);
}

// ---- matrices ----------------------------------------------------------------
__global__ void expand_lower_kernel(const double* __restrict__ tri, double* __restrict__ D, int n) {
    int j = blockIdx.x * blockDim.x + threadIdx.x;
    int i = blockIdx.y;
    if (j > i || i >= n) return;
    double v = (i == j) ? 0.0 : tri[(size_t)i * (i - 1) / 2 + j];
    D[(size_t)i * n + j] = v;
    D[(size_t)j * n + i] = v;
}
__global__ void symmetrize_kernel(double* __restrict__ D, int n) {
    // fillDismatrix (src/neighborJoining.cu:20-32): lower triangle wins, diagonal zero
    int j = blockIdx.x * blockDim.x + threadIdx.x;
    int i = blockIdx.y;
    if (j > i || i >= n) return;
    if (i == j) { D[(size_t)i * n + i] = 0.0; return; }
    D[(size_t)j * n + i] = D[(size_t)i * n + j];
}

// tiled transpose of rows [r0, r1) below the diagonal into the columns above it
__global__ void mirror_rows_kernel(double* __restrict__ D, int n, int r0, int r1) {
    __shared__ double tile[32][33];
    const int i0 = r0 + blockIdx.y * 32, j0 = blockIdx.x * 32;
    if (j0 > i0 + 31) return;                       // tile entirely above the diagonal
    const int tx = threadIdx.x, ty = threadIdx.y;   // 32 x 8
    for (int k = ty; k < 32; k += 8) {
        const int i = i0 + k, j = j0 + tx;
        tile[k][tx] = (i < r1 && j < i) ? D[(size_t)i * n + j] : 0.0;
    }
    __syncthreads();
    for (int k = ty; k < 32; k += 8) {
        const int j = j0 + k, i = i0 + tx;          // write D[j][i], coalesced along i
        if (i < r1 && j < i) D[(size_t)j * n + i] = tile[tx][k];
    }
}

int dipb_matrix_mirror_rows(dipb_matrix* m, int r0, int r1) {
    if (!m || r0 < 0 || r1 > m->n || r0 > r1) { set_error("dipb_matrix_mirror_rows: bad argument"); return DIPB_E_ARG; }
    if (r0 == r1) return 0;
    dipb_ctx* c = m->ctx;
    DIPB_CUDA(cudaSetDevice(c->device));
    int rc = timer_begin(c);
    if (rc) return rc;
    dim3 grid((r1 + 31) / 32, (r1 - r0 + 31) / 32), block(32, 8);
    mirror_rows_kernel<<<grid, block, 0, c->stream>>>(m->d, m->n, r0, r1);
    DIPB_KERNEL_CHECK(c);
    return timer_end(c, DIPB_T_MSA_DIST);
}

int dipb_matrix_from_host(dipb_ctx* c, const double* h, int n, int full, dipb_matrix** out) {
    if (!c || !h || !out || n < 2) { set_error("dipb_matrix_from_host: bad argument"); return DIPB_E_ARG; }
    DIPB_CUDA(cudaSetDevice(c->device));
    dipb_matrix* M = new dipb_matrix();
    M->ctx = c;
    ctx_retain(c);
    M->n = n;
    size_t bytes = (size_t)n * n * sizeof(double);
    cudaError_t e = pool_alloc(c, (void**)&M->d, bytes);
    if (e != cudaSuccess) { set_error("dipb_matrix_from_host: allocation of %zu bytes failed: %s", bytes, cudaGetErrorString(e)); M->d = nullptr; dipb_matrix_free(M); return DIPB_E_NOMEM; }
    dim3 grid((n + 255) / 256, n);
    double* tri = nullptr;
    auto body = [&]() -> int {
        if (full) {
            DIPB_CUDA(cudaMemcpyAsync(M->d, h, bytes, cudaMemcpyHostToDevice, c->stream));
            symmetrize_kernel<<<grid, 256, 0, c->stream>>>(M->d, n);
            DIPB_KERNEL_CHECK(c);
        } else {
            size_t tb = (size_t)n * (n - 1) / 2 * sizeof(double);
            DIPB_CUDA(cudaMalloc(&tri, tb));
            DIPB_CUDA(cudaMemcpyAsync(tri, h, tb, cudaMemcpyHostToDevice, c->stream));
            expand_lower_kernel<<<grid, 256, 0, c->stream>>>(tri, M->d, n);
            DIPB_KERNEL_CHECK(c);
        }
        DIPB_CUDA(cudaStreamSynchronize(c->stream));
        return 0;
    };
    const int rc = body();
    if (tri) { cudaStreamSynchronize(c->stream); cudaFree(tri); }
    if (rc) { dipb_matrix_free(M); return rc; }
    *out = M;
    return 0;
}

int dipb_matrix_n(const dipb_matrix* m) { return m ? m->n : 0; }
double* dipb_matrix_device_ptr(dipb_matrix* m) { return m ? m->d : nullptr; }
int dipb_matrix_to_host(dipb_matrix* m, double* h_out) {
    if (!m || !h_out) { set_error("dipb_matrix_to_host: bad argument"); return DIPB_E_ARG; }
    DIPB_CUDA(cudaSetDevice(m->ctx->device));
    DIPB_CUDA(cudaStreamSynchronize(m->ctx->stream));
    DIPB_CUDA(cudaMemcpy(h_out, m->d, (size_t)m->n * m->n * sizeof(double), cudaMemcpyDeviceToHost));
    return 0;
}
void dipb_matrix_free(dipb_matrix* m) {
    if (!m) return;
    cudaSetDevice(m->ctx->device);
    pool_free(m->ctx, m->d);
    ctx_release(m->ctx);
    delete m;
}

// ---- NJ ------------------------------------------------------------------------
int dipb_nj(dipb_matrix* m, int algo, int32_t* child0, int32_t* child1, double* len0, double* len1) {
    if (!m || !child0 || !child1 || !len0 || !len1) { set_error("dipb_nj: bad argument"); return DIPB_E_ARG; }
    if (algo < 0 || algo > DIPB_NJ_CLUSTER) { set_error("dipb_nj: unknown algorithm %d", algo); return DIPB_E_ARG; }
    DIPB_CUDA(cudaSetDevice(m->ctx->device));
    return nj_run(m, algo, child0, child1, len0, len1);
}
int dipb_nj_stats(dipb_ctx* c, uint64_t* rows_scanned, uint64_t* bytes_scanned, uint64_t* iterations) {
    if (!c) return DIPB_E_ARG;
    if (rows_scanned) *rows_scanned = c->nj_rows_scanned;
    if (bytes_scanned) *bytes_scanned = c->nj_bytes_scanned;
    if (iterations) *iterations = c->nj_iterations;
    return 0;
}

}  // extern "C"
