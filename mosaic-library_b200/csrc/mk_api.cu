// mk_api.cu -- version, error reporting and launch accounting of libmosaic_b200.
#include "mk_common.cuh"

#include <stdarg.h>

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
