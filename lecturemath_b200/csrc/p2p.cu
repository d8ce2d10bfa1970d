// p2p.cu -- the multi-GPU hand-off of the temporal-matching state over NVLink peer memory, without NCCL kernels:
// one process per GPU; every rank owns a mailbox (receive buffer + two 32-bit flags) that its ring neighbours map
// through CUDA IPC.  The sender's export kernels store the active unique-CC set STRAIGHT into the successor's
// mailbox (P2P stores over NVLink), then a stream memory operation (cuStreamWriteValue32) publishes "chunk n is
// there"; the receiver's stream blocks on cuStreamWaitValue32 until then, imports, and acknowledges the same way.
// No SM is occupied while a rank waits (NCCL's send/recv kernels spin on an SM until the peer arrives, which costs
// the persistent conv kernels an SM -- or a whole CTA pair -- for most of a step).
//
// This is plumbing of row (e) of the scope table (frame shards + one ordered exchange), not a
// reference function.
#include <cuda.h>

#include "am_common.cuh"
#include "../../include/accessmath_b200.h"

typedef CUresult (*PFN_writeValue32)(CUstream, CUdeviceptr, cuuint32_t, unsigned int);
typedef CUresult (*PFN_waitValue32)(CUstream, CUdeviceptr, cuuint32_t, unsigned int);

static void* driver_fn(const char* name) {
    void* sym = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint(name, &sym, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) return nullptr;
    return sym;
}

// bytes of device memory (cudaMalloc: legacy IPC cannot export stream-ordered / VMM allocations), zero-filled
extern "C" void* am_p2p_alloc(long long bytes) {
    void* p = nullptr;
    if (bytes <= 0 || cudaMalloc(&p, (size_t)bytes) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    if (cudaMemset(p, 0, (size_t)bytes) != cudaSuccess || cudaDeviceSynchronize() != cudaSuccess) { cudaFree(p); return nullptr; }
    return p;
}
extern "C" int am_p2p_free(void* d_ptr) {
    if (d_ptr) AM_CUDA(cudaFree(d_ptr));
    return AM_OK;
}
// h_handle: AM_P2P_HANDLE_BYTES (64) bytes the owner sends to its neighbours (any host channel)
extern "C" int am_p2p_export_handle(void* d_ptr, void* h_handle) {
    if (!d_ptr || !h_handle) return AM_ERR_ARG;
    cudaIpcMemHandle_t h;
    AM_CUDA(cudaIpcGetMemHandle(&h, d_ptr));
    static_assert(sizeof(h) == AM_P2P_HANDLE_BYTES, "cudaIpcMemHandle_t size");
    memcpy(h_handle, &h, sizeof(h));
    return AM_OK;
}
// maps a neighbour's mailbox into this process (enables peer access on first use)
extern "C" void* am_p2p_open_handle(const void* h_handle) {
    if (!h_handle) return nullptr;
    cudaIpcMemHandle_t h;
    memcpy(&h, h_handle, sizeof(h));
    void* p = nullptr;
    cudaError_t e = cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess);
    if (e != cudaSuccess) {
        fprintf(stderr, "[accessmath_b200] cudaIpcOpenMemHandle failed: %s\n", cudaGetErrorString(e));
        cudaGetLastError();
        return nullptr;
    }
    return p;
}
extern "C" int am_p2p_close_handle(void* d_peer_ptr) {
    if (d_peer_ptr) AM_CUDA(cudaIpcCloseMemHandle(d_peer_ptr));
    return AM_OK;
}
// stream-ordered 32-bit store (after everything enqueued before it, with a memory barrier): local or peer address
extern "C" int am_stream_write32(void* d_flag, unsigned value, void* stream) {
    static PFN_writeValue32 fn = (PFN_writeValue32)driver_fn("cuStreamWriteValue32");
    if (!fn || !d_flag) return AM_ERR_CUDA;
    CUresult r = fn((CUstream)stream, (CUdeviceptr)d_flag, value, CU_STREAM_WRITE_VALUE_DEFAULT);
    if (r != CUDA_SUCCESS) { fprintf(stderr, "[accessmath_b200] cuStreamWriteValue32 failed (%d)\n", (int)r); return AM_ERR_CUDA; }
    return AM_OK;
}
// the stream waits (no SM involved) until *(int32*)d_flag - value >= 0, i.e. the flag reached `value` (wrap-around safe)
extern "C" int am_stream_wait_geq32(void* d_flag, unsigned value, void* stream) {
    static PFN_waitValue32 fn = (PFN_waitValue32)driver_fn("cuStreamWaitValue32");
    if (!fn || !d_flag) return AM_ERR_CUDA;
    CUresult r = fn((CUstream)stream, (CUdeviceptr)d_flag, value, CU_STREAM_WAIT_VALUE_GEQ);
    if (r != CUDA_SUCCESS) { fprintf(stderr, "[accessmath_b200] cuStreamWaitValue32 failed (%d)\n", (int)r); return AM_ERR_CUDA; }
    return AM_OK;
}
