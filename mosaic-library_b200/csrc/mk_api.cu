// mk_api.cu -- version, error reporting and launch accounting of libmosaic_b200.
#include "mk_common.cuh"

#include <cuda.h>

#include <stdarg.h>
#include <string.h>

namespace mk {

static thread_local char t_err[512] = "";
std::atomic<uint64_t> g_launches{0};

void set_error(const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(t_err, sizeof t_err, fmt, ap);
    va_end(ap);
}

}  // namespace mk

extern "C" int mk_abi_version(void) { return MK_ABI_VERSION; }
extern "C" const char *mk_version_string(void) { return "mosaic_b200 0.1 (sm_100a)"; }
extern "C" const char *mk_last_error(void) { return mk::t_err; }
extern "C" uint64_t mk_launch_count(void) { return mk::g_launches.load(); }

extern "C" int mk_upload_2d(void *dst_dev, size_t dst_pitch, const void *src_host, size_t src_pitch, size_t row_bytes, size_t rows,
                            void *stream)
{
    if (!dst_dev || !src_host) { mk::set_error("mk_upload_2d: null pointer"); return MK_ERR_ARG; }
    if (rows == 0 || row_bytes == 0) return MK_OK;
    if (dst_pitch == row_bytes && src_pitch == row_bytes)   // dense on both sides: one linear copy (faster than per-row DMA descriptors)
        MK_CUDA_TRY(cudaMemcpyAsync(dst_dev, src_host, row_bytes * rows, cudaMemcpyHostToDevice, (cudaStream_t)stream));
    else
        MK_CUDA_TRY(cudaMemcpy2DAsync(dst_dev, dst_pitch, src_host, src_pitch, row_bytes, rows, cudaMemcpyHostToDevice, (cudaStream_t)stream));
    return MK_OK;
}

extern "C" int mk_download_2d(void *dst_host, size_t dst_pitch, const void *src_dev, size_t src_pitch, size_t row_bytes, size_t rows,
                              void *stream)
{
    if (!dst_host || !src_dev) { mk::set_error("mk_download_2d: null pointer"); return MK_ERR_ARG; }
    if (rows == 0 || row_bytes == 0) return MK_OK;
    if (dst_pitch == row_bytes && src_pitch == row_bytes)
        MK_CUDA_TRY(cudaMemcpyAsync(dst_host, src_dev, row_bytes * rows, cudaMemcpyDeviceToHost, (cudaStream_t)stream));
    else
        MK_CUDA_TRY(cudaMemcpy2DAsync(dst_host, dst_pitch, src_dev, src_pitch, row_bytes, rows, cudaMemcpyDeviceToHost, (cudaStream_t)stream));
    return MK_OK;
}

extern "C" int mk_copy_peer(void *dst_dev, const void *src_dev, size_t bytes, void *stream)
{
    if (!dst_dev || !src_dev) { mk::set_error("mk_copy_peer: null pointer"); return MK_ERR_ARG; }
    if (bytes == 0) return MK_OK;
    MK_CUDA_TRY(cudaMemcpyAsync(dst_dev, src_dev, bytes, cudaMemcpyDefault, (cudaStream_t)stream));
    return MK_OK;
}

extern "C" void *mk_event_create(void)
{
    cudaEvent_t e = nullptr;
    if (cudaEventCreateWithFlags(&e, cudaEventDisableTiming) != cudaSuccess) { mk::set_error("cudaEventCreateWithFlags failed"); return nullptr; }
    return e;
}

extern "C" void mk_event_destroy(void *event)
{
    if (event) cudaEventDestroy(static_cast<cudaEvent_t>(event));
}

extern "C" int mk_stream_follow(void *event, void *src_stream, void *dst_stream)
{
    if (!event) { mk::set_error("mk_stream_follow: null event"); return MK_ERR_ARG; }
    if (src_stream != MK_STREAM_NONE) MK_CUDA_TRY(cudaEventRecord(static_cast<cudaEvent_t>(event), static_cast<cudaStream_t>(src_stream)));
    if (dst_stream != MK_STREAM_NONE) MK_CUDA_TRY(cudaStreamWaitEvent(static_cast<cudaStream_t>(dst_stream), static_cast<cudaEvent_t>(event), 0));
    return MK_OK;
}

// The driver API is reached through cudaGetDriverEntryPoint only: the library must load (and export its symbols) on
// machines without libcuda.so.1, e.g. the CPU-only build check.
namespace {
template <class Fn>
Fn driver_fn(const char *name)
{
    void *p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint(name, &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) return nullptr;
    return reinterpret_cast<Fn>(p);
}
struct DriverApi {
    CUresult (*CtxGetDevice)(CUdevice *) = nullptr;
    CUresult (*DeviceGetDevResource)(CUdevice, CUdevResource *, CUdevResourceType) = nullptr;
    CUresult (*DevSmResourceSplitByCount)(CUdevResource *, unsigned int *, const CUdevResource *, CUdevResource *, unsigned int, unsigned int) = nullptr;
    CUresult (*DevResourceGenerateDesc)(CUdevResourceDesc *, CUdevResource *, unsigned int) = nullptr;
    CUresult (*GreenCtxCreate)(CUgreenCtx *, CUdevResourceDesc, CUdevice, unsigned int) = nullptr;
    CUresult (*GreenCtxDestroy)(CUgreenCtx) = nullptr;
    CUresult (*GreenCtxStreamCreate)(CUstream *, CUgreenCtx, unsigned int, int) = nullptr;
    CUresult (*GreenCtxGetDevResource)(CUgreenCtx, CUdevResource *, CUdevResourceType) = nullptr;
    CUresult (*StreamGetGreenCtx)(CUstream, CUgreenCtx *) = nullptr;
    CUresult (*StreamDestroy)(CUstream) = nullptr;
    bool ok = false;
};
const DriverApi &driver()
{
    static const DriverApi api = [] {
        DriverApi d;
        d.CtxGetDevice = driver_fn<decltype(d.CtxGetDevice)>("cuCtxGetDevice");
        d.DeviceGetDevResource = driver_fn<decltype(d.DeviceGetDevResource)>("cuDeviceGetDevResource");
        d.DevSmResourceSplitByCount = driver_fn<decltype(d.DevSmResourceSplitByCount)>("cuDevSmResourceSplitByCount");
        d.DevResourceGenerateDesc = driver_fn<decltype(d.DevResourceGenerateDesc)>("cuDevResourceGenerateDesc");
        d.GreenCtxCreate = driver_fn<decltype(d.GreenCtxCreate)>("cuGreenCtxCreate");
        d.GreenCtxDestroy = driver_fn<decltype(d.GreenCtxDestroy)>("cuGreenCtxDestroy");
        d.GreenCtxStreamCreate = driver_fn<decltype(d.GreenCtxStreamCreate)>("cuGreenCtxStreamCreate");
        d.GreenCtxGetDevResource = driver_fn<decltype(d.GreenCtxGetDevResource)>("cuGreenCtxGetDevResource");
        d.StreamGetGreenCtx = driver_fn<decltype(d.StreamGetGreenCtx)>("cuStreamGetGreenCtx");
        d.StreamDestroy = driver_fn<decltype(d.StreamDestroy)>("cuStreamDestroy");
        d.ok = d.CtxGetDevice && d.DeviceGetDevResource && d.DevSmResourceSplitByCount && d.DevResourceGenerateDesc && d.GreenCtxCreate &&
               d.GreenCtxDestroy && d.GreenCtxStreamCreate && d.GreenCtxGetDevResource && d.StreamGetGreenCtx && d.StreamDestroy;
        return d;
    }();
    return api;
}
struct SmPartition {
    CUgreenCtx ctx[2] = { nullptr, nullptr };
    CUstream stream[2] = { nullptr, nullptr };
};
std::atomic<unsigned> g_partition_epoch{0};   // bumped whenever a partition dies: a recycled stream handle must not hit the cache
}  // namespace

extern "C" int mk_sm_partition_create(int min_sms_first, void **partition, void **stream_first, void **stream_rest, int *sms_first,
                                      int *sms_rest)
{
    if (!partition || !stream_first || !stream_rest || min_sms_first <= 0) { mk::set_error("mk_sm_partition_create: bad argument"); return MK_ERR_ARG; }
    cudaFree(nullptr);                                   // make sure the primary context exists and is current
    const DriverApi &D = driver();
    if (!D.ok) { mk::set_error("mk_sm_partition_create: this driver has no green contexts"); return MK_ERR_CUDA; }
    CUdevice dev;
    CUdevResource all, first, rest;
    unsigned int groups = 1;
#define MK_CU_TRY(expr)                                                                                    \
    do {                                                                                                   \
        CUresult r_ = (expr);                                                                              \
        if (r_ != CUDA_SUCCESS) {                                                                          \
            mk::set_error("%s: driver error %d", #expr, (int)r_);                                          \
            if (P) mk_sm_partition_destroy(P);                                                             \
            return MK_ERR_CUDA;                                                                            \
        }                                                                                                  \
    } while (0)
    SmPartition *P = nullptr;
    MK_CU_TRY(D.CtxGetDevice(&dev));
    MK_CU_TRY(D.DeviceGetDevResource(dev, &all, CU_DEV_RESOURCE_TYPE_SM));
    MK_CU_TRY(D.DevSmResourceSplitByCount(&first, &groups, &all, &rest, 0, (unsigned int)min_sms_first));
    if (groups < 1 || rest.sm.smCount == 0) { mk::set_error("mk_sm_partition_create: cannot split %u SMs at %d", all.sm.smCount, min_sms_first); return MK_ERR_ARG; }
    P = new SmPartition();
    CUdevResource parts[2] = { first, rest };
    for (int i = 0; i < 2; ++i) {
        CUdevResourceDesc desc;
        MK_CU_TRY(D.DevResourceGenerateDesc(&desc, &parts[i], 1));
        MK_CU_TRY(D.GreenCtxCreate(&P->ctx[i], desc, dev, CU_GREEN_CTX_DEFAULT_STREAM));
        MK_CU_TRY(D.GreenCtxStreamCreate(&P->stream[i], P->ctx[i], CU_STREAM_NON_BLOCKING, 0));
    }
#undef MK_CU_TRY
    *partition = P;
    *stream_first = P->stream[0];
    *stream_rest = P->stream[1];
    if (sms_first) *sms_first = (int)first.sm.smCount;
    if (sms_rest) *sms_rest = (int)rest.sm.smCount;
    return MK_OK;
}

extern "C" void mk_sm_partition_destroy(void *partition)
{
    SmPartition *P = static_cast<SmPartition *>(partition);
    if (!P) return;
    for (int i = 0; i < 2; ++i) {
        if (P->stream[i]) driver().StreamDestroy(P->stream[i]);
        if (P->ctx[i]) driver().GreenCtxDestroy(P->ctx[i]);
    }
    g_partition_epoch.fetch_add(1, std::memory_order_release);
    delete P;
}

namespace mk {
// SMs available to work launched on `stream`: the size of its green-context partition, or the whole device.
int stream_sm_count(cudaStream_t stream)
{
    // one-entry cache keyed by (stream handle, current device, partition epoch): a handle can be recycled after
    // mk_sm_partition_destroy, and the NULL stream means a different device after cudaSetDevice
    static thread_local cudaStream_t last_stream = reinterpret_cast<cudaStream_t>(~(uintptr_t)0);
    static thread_local int last_count = 0, last_dev = -1;
    static thread_local unsigned last_epoch = 0;
    int cur_dev = 0;
    cudaGetDevice(&cur_dev);
    const unsigned epoch = g_partition_epoch.load(std::memory_order_acquire);
    if (stream == last_stream && cur_dev == last_dev && epoch == last_epoch && last_count > 0) return last_count;
    int count = 0;
    CUgreenCtx g = nullptr;
    if (stream && driver().ok && driver().StreamGetGreenCtx(reinterpret_cast<CUstream>(stream), &g) == CUDA_SUCCESS && g) {
        CUdevResource res;
        if (driver().GreenCtxGetDevResource(g, &res, CU_DEV_RESOURCE_TYPE_SM) == CUDA_SUCCESS) count = (int)res.sm.smCount;
    }
    if (count <= 0) cudaDeviceGetAttribute(&count, cudaDevAttrMultiProcessorCount, cur_dev);
    last_stream = stream; last_count = count; last_dev = cur_dev; last_epoch = epoch;
    return count;
}
}  // namespace mk

// ---------------------------------------------------------------------------------------------
// Peer memory over NVLink / NVSwitch for the tile-sharded canvas (one process per GPU): export a device allocation, map it in
// the other ranks, hand the mapped pointer to mk_mapper_update as `src`.
// ---------------------------------------------------------------------------------------------
typedef CUresult (*GetAddressRangeFn)(CUdeviceptr *, size_t *, CUdeviceptr);

extern "C" int mk_ipc_export(const void *dev_ptr, uint8_t *handle_out, uint64_t *offset_out)
{
    if (!dev_ptr || !handle_out || !offset_out) { mk::set_error("mk_ipc_export: null pointer"); return MK_ERR_ARG; }
    static GetAddressRangeFn fn = [] {
        void *q = nullptr;
        cudaDriverEntryPointQueryResult qr;
        if (cudaGetDriverEntryPoint("cuMemGetAddressRange", &q, cudaEnableDefault, &qr) != cudaSuccess || qr != cudaDriverEntryPointSuccess) q = nullptr;
        return reinterpret_cast<GetAddressRangeFn>(q);
    }();
    if (!fn) { mk::set_error("mk_ipc_export: cuMemGetAddressRange not available"); return MK_ERR_CUDA; }
    CUdeviceptr base = 0;
    size_t size = 0;
    if (fn(&base, &size, (CUdeviceptr)dev_ptr) != CUDA_SUCCESS) { mk::set_error("mk_ipc_export: cuMemGetAddressRange failed (not a device allocation?)"); return MK_ERR_CUDA; }
    cudaIpcMemHandle_t h;
    MK_CUDA_TRY(cudaIpcGetMemHandle(&h, reinterpret_cast<void *>(base)));
    static_assert(sizeof(h) == 64, "cudaIpcMemHandle_t is 64 bytes");
    memcpy(handle_out, &h, 64);
    *offset_out = (uint64_t)((CUdeviceptr)dev_ptr - base);
    return MK_OK;
}

extern "C" int mk_ipc_open(const uint8_t *handle, void **base_out)
{
    if (!handle || !base_out) { mk::set_error("mk_ipc_open: null pointer"); return MK_ERR_ARG; }
    cudaIpcMemHandle_t h;
    memcpy(&h, handle, 64);
    MK_CUDA_TRY(cudaIpcOpenMemHandle(base_out, h, cudaIpcMemLazyEnablePeerAccess));
    return MK_OK;
}

extern "C" int mk_ipc_close(void *base)
{
    if (!base) return MK_OK;
    MK_CUDA_TRY(cudaIpcCloseMemHandle(base));
    return MK_OK;
}
