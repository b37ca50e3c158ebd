// Runtime part of the C ABI: handles, error reporting, workspace, host-buffer entry point of K1.
#include <stdarg.h>
#include <stdlib.h>
#include "lr_common.cuh"

static thread_local char g_err[1024] = "";

void lr_set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

extern "C" const char* lr_last_error(void) { return g_err; }
extern "C" int lr_abi_version(void) { return LR_ABI_VERSION; }

extern "C" int lr_create(int device, lr_handle_t* out) {
    LR_REQUIRE(out != nullptr, "lr_create: null output pointer");
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0) {
        lr_set_error("lr_create: no CUDA device (%s); literate_b200 has no CPU path", e == cudaSuccess ? "count = 0" : cudaGetErrorString(e));
        return LR_ERR_CUDA;
    }
    LR_REQUIRE(device >= 0 && device < n, "lr_create: device %d out of range (0..%d)", device, n - 1);
    LR_CUDA(cudaSetDevice(device));
    cudaDeviceProp prop;
    LR_CUDA(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10) {
        lr_set_error("lr_create: device %d is sm_%d%d; this library is built for sm_100a (B200) only", device, prop.major, prop.minor);
        return LR_ERR_UNSUPPORTED;
    }
    lr_handle_t h = new lr_handle_s();
    memset(h, 0, sizeof(*h));
    h->device = device;
    h->sm_count = prop.multiProcessorCount;
    h->max_smem_optin = (int)prop.sharedMemPerBlockOptin;
    // Datasets and chains are allocated stream-ordered (cudaMallocAsync).  Keep freed blocks in the device's pool instead of
    // returning them to the driver at every synchronisation (measured: 75-290 ms of unmapping per device-wide sync).
    {
        cudaMemPool_t pool;
        LR_CUDA(cudaDeviceGetDefaultMemPool(&pool, device));
        unsigned long long keep = ~0ull;
        LR_CUDA(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep));
    }
    LR_CUDA(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
    LR_CUDA(cudaStreamCreateWithFlags(&h->copy_stream, cudaStreamNonBlocking));
    for (int i = 0; i < 2 * LR_NSTAGE; ++i) LR_CUDA(cudaEventCreateWithFlags(&h->ev[i], cudaEventDisableTiming));
    LR_CUDA(cudaEventCreateWithFlags(&h->ev_order, cudaEventDisableTiming));
    LR_CUDA(cudaHostAlloc((void**)&h->k1_hint, sizeof(int), cudaHostAllocMapped | cudaHostAllocPortable));
    *h->k1_hint = 0;
    *out = h;
    return LR_OK;
}

extern "C" int lr_destroy(lr_handle_t h) {
    if (!h) return LR_OK;
    cudaSetDevice(h->device);
    cudaStreamSynchronize(h->stream);
    cudaStreamSynchronize(h->copy_stream);
    if (h->ws_used || h->k1_hint_used) cudaDeviceSynchronize();        // the last user of the workspace / the K1 hint may have been a caller stream
    cudaFree(h->ws);
    cudaFreeHost(h->k1_hint);
    for (int i = 0; i < LR_NSTAGE; ++i) cudaFree(h->stage[i]);
    for (int i = 0; i < 2 * LR_NSTAGE; ++i) cudaEventDestroy(h->ev[i]);
    cudaEventDestroy(h->ev_order);
    cudaStreamDestroy(h->stream);
    cudaStreamDestroy(h->copy_stream);
    delete h;
    return LR_OK;
}

extern "C" int lr_info(lr_handle_t h, int32_t* sm_count, int64_t* kernel_launches, int32_t* device) {
    LR_REQUIRE(h != nullptr, "lr_info: null handle");
    if (sm_count) *sm_count = h->sm_count;
    if (kernel_launches) *kernel_launches = h->launches;
    if (device) *device = h->device;
    return LR_OK;
}

extern "C" int lr_sync(lr_handle_t h) {
    LR_REQUIRE(h != nullptr, "lr_sync: null handle");
    LR_CUDA(cudaSetDevice(h->device));
    LR_CUDA(cudaStreamSynchronize(h->copy_stream));
    LR_CUDA(cudaStreamSynchronize(h->stream));
    return LR_OK;
}

int lr_order(lr_handle_t h, lr_last_stream* last, cudaStream_t now) {
    if (last->valid && last->s != now) {
        // everything submitted to the previous stream so far happens before whatever `now` gets from here on
        cudaError_t e = cudaEventRecord(h->ev_order, last->s);
        if (e == cudaSuccess) e = cudaStreamWaitEvent(now, h->ev_order, 0);
        if (e != cudaSuccess) {                 // e.g. the caller destroyed that stream: fall back to a full barrier
            cudaGetLastError();
            LR_CUDA(cudaDeviceSynchronize());
        }
    }
    last->s = now; last->valid = 1;
    return LR_OK;
}

int lr_ws_acquire(lr_handle_t h, size_t bytes, cudaStream_t stream) {
    if (bytes > h->ws_bytes) {
        LR_CUDA(cudaDeviceSynchronize());       // growing is rare; nobody may still be using the old block
        if (h->ws) LR_CUDA(cudaFree(h->ws));
        h->ws = nullptr; h->ws_bytes = 0;
        size_t want = (bytes + ((size_t)1 << 20) - 1) & ~(((size_t)1 << 20) - 1);
        cudaError_t e = cudaMalloc(&h->ws, want);
        if (e != cudaSuccess) {
            lr_set_error("workspace allocation of %zu bytes failed: %s", want, cudaGetErrorString(e));
            return LR_ERR_NOMEM;
        }
        h->ws_bytes = want;
        h->ws_used = 0;
    }
    lr_last_stream last = {h->ws_stream, h->ws_used};
    int rc = lr_order(h, &last, stream);
    h->ws_stream = stream; h->ws_used = 1;
    return rc;
}

static int stage_reserve(lr_handle_t h, size_t bytes) {
    if (bytes <= h->stage_bytes) return LR_OK;
    LR_CUDA(cudaStreamSynchronize(h->stream));
    LR_CUDA(cudaStreamSynchronize(h->copy_stream));
    for (int i = 0; i < LR_NSTAGE; ++i) {
        if (h->stage[i]) LR_CUDA(cudaFree(h->stage[i]));
        h->stage[i] = nullptr;
    }
    h->stage_bytes = 0;
    for (int i = 0; i < LR_NSTAGE; ++i) {
        cudaError_t e = cudaMalloc(&h->stage[i], bytes);
        if (e != cudaSuccess) {
            lr_set_error("staging allocation of %zu bytes failed: %s", bytes, cudaGetErrorString(e));
            return LR_ERR_NOMEM;
        }
    }
    h->stage_bytes = bytes;
    return LR_OK;
}

// Host buffers in, host buffers out.  Replicates are copied in batches on the copy stream into
// two device staging buffers while the previous batch is being binned on the compute stream.
// Pass pinned host memory for the copies to be truly asynchronous.
// elem = 8: fp64 times, `frac` = fe_ref;  elem = 4: int32 years, `frac` = death_jitter (half the bytes over PCIe).
static int bin_stats_host_impl(lr_handle_t h, const void* h_ts, const void* h_te, int elem, int64_t n, int64_t ld,
                               int32_t n_rep, int64_t first_bin, int32_t n_bins, double frac,
                               int32_t dead_only, double end_time,
                               int64_t* h_sp, int64_t* h_ex, double* h_br, const char* who) {
    LR_REQUIRE(h != nullptr, "%s: null handle", who);
    LR_REQUIRE(n >= 0 && n_rep >= 1 && ld >= n && n_bins >= 1, "%s: bad sizes", who);
    LR_REQUIRE(h_sp && h_ex && h_br && (n == 0 || (h_ts && h_te)), "%s: null pointer", who);
    LR_CUDA(cudaSetDevice(h->device));
    const double fe_ref = elem == 8 ? frac : lr_fe_ref_of_jitter(frac);
    const size_t stride = (size_t)lr_acc_stride(n_bins);
    const size_t acc_bytes = (size_t)n_rep * LR_ACC_ROWS * stride * sizeof(int64_t);
    const size_t out_cnt = (size_t)n_rep * n_bins;
    const size_t acc_pad = (acc_bytes + 255) & ~(size_t)255;
    int rc = lr_ws_acquire(h, acc_pad + out_cnt * 24, h->stream);
    if (rc != LR_OK) return rc;
    int64_t* d_acc = (int64_t*)h->ws;
    int64_t* d_sp = (int64_t*)((char*)h->ws + acc_pad);
    int64_t* d_ex = d_sp + out_cnt;
    double* d_br = (double*)(d_ex + out_cnt);
    LR_CUDA(cudaMemsetAsync(d_acc, 0, acc_bytes, h->stream));

    if (n > 0) {
        const int64_t align = 16 / elem;                            // row pitch that keeps 128-bit loads aligned
        const int64_t ldp = (n + align - 1) / align * align;
        static int stage_mb = 0;                                    // development knob: LR_STAGE_MB (default 64 MB of (ts, te) per batch)
        if (!stage_mb) { const char* e = getenv("LR_STAGE_MB"); stage_mb = e && atoi(e) > 0 ? atoi(e) : 64; }
        int64_t per_batch = ((int64_t)stage_mb << 20) / (ldp * 2 * elem);
        if (per_batch < 1) per_batch = 1;
        if (per_batch > n_rep) per_batch = n_rep;
        const size_t half = (size_t)per_batch * ldp * elem;
        rc = stage_reserve(h, 2 * half);
        if (rc != LR_OK) return rc;
        int batch = 0;
        // development aid (LR_DEBUG_TIMELINE=1): time stamps of every batch's copies and kernel on their streams
        static int dbg = -1;
        if (dbg < 0) { const char* e = getenv("LR_DEBUG_TIMELINE"); dbg = e && atoi(e) ? 1 : 0; }
        const int max_dbg = 512;
        cudaEvent_t* te = nullptr;
        if (dbg) {
            te = new cudaEvent_t[4 * max_dbg];
            for (int i = 0; i < 4 * max_dbg; ++i) cudaEventCreate(&te[i]);
        }
        for (int64_t r0 = 0; r0 < n_rep; r0 += per_batch, ++batch) {
            const int buf = batch % LR_NSTAGE;
            const int64_t nr = (n_rep - r0 < per_batch) ? (n_rep - r0) : per_batch;
            char* s_ts = (char*)h->stage[buf];
            char* s_te = (char*)h->stage[buf] + half;
            if (batch >= LR_NSTAGE) LR_CUDA(cudaStreamWaitEvent(h->copy_stream, h->ev[LR_NSTAGE + buf], 0));     // the kernel that read this buffer last is done
            if (dbg && batch < max_dbg) cudaEventRecord(te[4 * batch + 0], h->copy_stream);
            if (ld == ldp && ld == n) {      // rows back to back on both sides: one linear copy per array
                LR_CUDA(cudaMemcpyAsync(s_ts, (const char*)h_ts + r0 * ld * elem, (size_t)nr * n * elem, cudaMemcpyHostToDevice, h->copy_stream));
                LR_CUDA(cudaMemcpyAsync(s_te, (const char*)h_te + r0 * ld * elem, (size_t)nr * n * elem, cudaMemcpyHostToDevice, h->copy_stream));
            } else {
                LR_CUDA(cudaMemcpy2DAsync(s_ts, ldp * elem, (const char*)h_ts + r0 * ld * elem, ld * elem, n * elem, nr, cudaMemcpyHostToDevice, h->copy_stream));
                LR_CUDA(cudaMemcpy2DAsync(s_te, ldp * elem, (const char*)h_te + r0 * ld * elem, ld * elem, n * elem, nr, cudaMemcpyHostToDevice, h->copy_stream));
            }
            if (dbg && batch < max_dbg) cudaEventRecord(te[4 * batch + 1], h->copy_stream);
            LR_CUDA(cudaEventRecord(h->ev[buf], h->copy_stream));
            LR_CUDA(cudaStreamWaitEvent(h->stream, h->ev[buf], 0));
            if (dbg && batch < max_dbg) cudaEventRecord(te[4 * batch + 2], h->stream);
            int64_t* acc_r = d_acc + (size_t)r0 * LR_ACC_ROWS * stride;
            h->k1_general_only = 1;     // batches run beside the chains of the previous table and behind the host link: see lr_common.cuh
            rc = elem == 8 ? lr_bin_accumulate(h, (const double*)s_ts, (const double*)s_te, n, ldp, (int32_t)nr, first_bin, n_bins, fe_ref, dead_only,
                                               end_time, acc_r, h->stream)
                           : lr_bin_accumulate_i32(h, (const int32_t*)s_ts, (const int32_t*)s_te, n, ldp, (int32_t)nr, first_bin, n_bins, frac,
                                                   dead_only, end_time, acc_r, h->stream);
            h->k1_general_only = 0;
            if (rc != LR_OK) return rc;
            LR_CUDA(cudaEventRecord(h->ev[LR_NSTAGE + buf], h->stream));
            if (dbg && batch < max_dbg) cudaEventRecord(te[4 * batch + 3], h->stream);
        }
        if (dbg) {
            cudaStreamSynchronize(h->stream); cudaStreamSynchronize(h->copy_stream);
            const int nbt = batch < max_dbg ? batch : max_dbg;
            float copy_busy = 0, k_busy = 0, span = 0, k_max = 0, gap = 0;
            for (int b = 0; b < nbt; ++b) {
                float c = 0, k = 0;
                cudaEventElapsedTime(&c, te[4 * b], te[4 * b + 1]); cudaEventElapsedTime(&k, te[4 * b + 2], te[4 * b + 3]);
                copy_busy += c; k_busy += k; if (k > k_max) k_max = k;
                if (b > 0) { float g = 0; cudaEventElapsedTime(&g, te[4 * (b - 1) + 1], te[4 * b]); gap += g; }
            }
            cudaEventElapsedTime(&span, te[0], te[4 * (nbt - 1) + 3]);
            fprintf(stderr, "[lr timeline] %d batches: span %.2f ms, copies busy %.2f ms, idle between copies %.2f ms, kernels busy %.2f ms (max %.3f)\n",
                    nbt, span, copy_busy, gap, k_busy, k_max);
            for (int i = 0; i < 4 * max_dbg; ++i) cudaEventDestroy(te[i]);
            delete[] te;
        }
    }
    rc = lr_bin_finalize(h, d_acc, n_rep, n_bins, fe_ref, d_sp, d_ex, d_br, h->stream);
    if (rc != LR_OK) return rc;
    LR_CUDA(cudaMemcpyAsync(h_sp, d_sp, out_cnt * 8, cudaMemcpyDeviceToHost, h->stream));
    LR_CUDA(cudaMemcpyAsync(h_ex, d_ex, out_cnt * 8, cudaMemcpyDeviceToHost, h->stream));
    LR_CUDA(cudaMemcpyAsync(h_br, d_br, out_cnt * 8, cudaMemcpyDeviceToHost, h->stream));
    LR_CUDA(cudaStreamSynchronize(h->stream));
    return LR_OK;
}

extern "C" int lr_bin_stats_host(lr_handle_t h, const double* h_ts, const double* h_te, int64_t n, int64_t ld,
                                 int32_t n_rep, int64_t first_bin, int32_t n_bins, double fe_ref,
                                 int32_t dead_only, double end_time,
                                 int64_t* h_sp, int64_t* h_ex, double* h_br) {
    return bin_stats_host_impl(h, h_ts, h_te, 8, n, ld, n_rep, first_bin, n_bins, fe_ref, dead_only, end_time, h_sp, h_ex, h_br, "lr_bin_stats_host");
}

extern "C" int lr_bin_stats_host_i32(lr_handle_t h, const int32_t* h_ts, const int32_t* h_te, int64_t n, int64_t ld,
                                     int32_t n_rep, int64_t first_bin, int32_t n_bins, double death_jitter,
                                     int32_t dead_only, double end_time,
                                     int64_t* h_sp, int64_t* h_ex, double* h_br) {
    return bin_stats_host_impl(h, h_ts, h_te, 4, n, ld, n_rep, first_bin, n_bins, death_jitter, dead_only, end_time, h_sp, h_ex, h_br,
                               "lr_bin_stats_host_i32");
}
