// Host side of the tcgen05/TMA GEMM: tensor-map creation (driver entry point fetched at run time so the
// library has no link-time dependency on libcuda) and launch wrappers.
#define B2A_TC_GEMM_IMPL
#include "common.cuh"
#include "tc_gemm.cuh"

namespace b2a {
namespace tc {

static EncodeTiledFn encode_fn() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        B2A_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q));
        B2A_CHECK(p && q == cudaDriverEntryPointSuccess, B2A_ERR_CUDA, "cuTensorMapEncodeTiled is not available in this driver");
        fn = (EncodeTiledFn)p;
    }
    return fn;
}

CUtensorMap make_tmap_bf16(const void* base, long long rows, long long cols, int box_rows) {
    B2A_CHECK(cols % 8 == 0 && ((uintptr_t)base & 15) == 0, B2A_ERR_INVALID_INPUT, "TMA: tensor must be 16-byte aligned with cols % 8 == 0");
    CUtensorMap m;
    const cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    const cuuint64_t strides[1] = {(cuuint64_t)cols * 2};
    const cuuint32_t box[2] = {(cuuint32_t)BK, (cuuint32_t)box_rows};
    const cuuint32_t estr[2] = {1, 1};
    const CUresult r = encode_fn()(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    B2A_CHECK(r == CUDA_SUCCESS, B2A_ERR_CUDA, "cuTensorMapEncodeTiled failed (" + std::to_string((int)r) + ")");
    return m;
}

// 3-D fp16 tensor [d2][d1][d0] (d0 contiguous), box {b0, b1, 1}, 128-byte swizzle (b0 * 2 bytes must be 128): attn_tc.cuh's operands
CUtensorMap make_tmap_f16_3d(const void* base, long long d0, long long d1, long long d2, int b0, int b1) {
    B2A_CHECK(b0 * 2 == 128 && d0 % 8 == 0 && ((uintptr_t)base & 15) == 0, B2A_ERR_INVALID_INPUT, "TMA: bad 3-D fp16 tensor");
    CUtensorMap m;
    const cuuint64_t dims[3] = {(cuuint64_t)d0, (cuuint64_t)d1, (cuuint64_t)d2};
    const cuuint64_t strides[2] = {(cuuint64_t)d0 * 2, (cuuint64_t)d0 * (cuuint64_t)d1 * 2};
    const cuuint32_t box[3] = {(cuuint32_t)b0, (cuuint32_t)b1, 1};
    const cuuint32_t estr[3] = {1, 1, 1};
    const CUresult r = encode_fn()(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, const_cast<void*>(base), dims, strides, box, estr,
                                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    B2A_CHECK(r == CUDA_SUCCESS, B2A_ERR_CUDA, "cuTensorMapEncodeTiled (3-D fp16) failed (" + std::to_string((int)r) + ")");
    return m;
}

template <int BN>
void launch(const CUtensorMap& tmA, const CUtensorMap& tmB, const Args& a, int ctas, int n_tiles, cudaStream_t s) {
    launch_pdl(tc_gemm_kernel<BN>, dim3(ctas, n_tiles), dim3(Cfg<BN>::NTHREADS), Smem<BN>::bytes(a.stages), s, tmA, tmB, a);
}
template void launch<16>(const CUtensorMap&, const CUtensorMap&, const Args&, int, int, cudaStream_t);
template void launch<32>(const CUtensorMap&, const CUtensorMap&, const Args&, int, int, cudaStream_t);
template void launch<128>(const CUtensorMap&, const CUtensorMap&, const Args&, int, int, cudaStream_t);

void launch_splitk(const CUtensorMap& tmA, const CUtensorMap& tmB, const SplitArgs& a, int m_tiles, int cluster, cudaStream_t s) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3((unsigned)(m_tiles * cluster)); cfg.blockDim = dim3(THREADS);
    cfg.dynamicSmemBytes = SmemSplit::bytes(a.stages, cluster); cfg.stream = s;
    cudaLaunchAttribute at[2];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = pdl_enabled() ? 1 : 0;
    at[1].id = cudaLaunchAttributeClusterDimension;
    at[1].val.clusterDim.x = (unsigned)cluster; at[1].val.clusterDim.y = 1; at[1].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 2;
    B2A_CUDA(cudaLaunchKernelEx(&cfg, tc_gemm_splitk_kernel, tmA, tmB, a));
    count_launch();
}

void set_attributes() {
    B2A_CUDA(cudaFuncSetAttribute(tc_gemm_splitk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    B2A_CUDA(cudaFuncSetAttribute(tc_gemm_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    B2A_CUDA(cudaFuncSetAttribute(tc_gemm_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    B2A_CUDA(cudaFuncSetAttribute(tc_gemm_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
}

}  // namespace tc
}  // namespace b2a

// Standalone entry used by tests/test_gpu_tc_gemm.py: out[N, M] = X[N, K] * W[M, K]^T through the tcgen05 path.
// All pointers are DEVICE pointers.  epi: 0 store fp32, 2 SwiGLU (bf16 out [N(+lo), M/2]), 3 store bf16;
// split != 0 allows stream-K partial tiles (atomic adds into the zero-initialised fp32 output).
extern "C" int32_t b2a_tc_gemm_test(const void* W, const void* X, void* out, int32_t M, int32_t N, int32_t K, int32_t bn,
                                    int32_t epi, int32_t split, int32_t hilo, int32_t ctas, void* stream) {
    using namespace b2a;
    using namespace b2a::tc;
    return guarded([&] {
        B2A_CHECK(W && X && out && (bn == 16 || bn == 32 || bn == 128) && K % BK == 0, B2A_ERR_INVALID_INPUT, "b2a_tc_gemm_test: bad argument");
        require_device(0);
        set_attributes();
        const int x_rows = hilo ? bn : N;
        CUtensorMap ta = make_tmap_bf16(W, M, K, BM), tb = make_tmap_bf16(X, x_rows, K, bn);
        Args a{};
        a.out_f32 = (float*)out; a.out_bf16 = (__nv_bfloat16*)out; a.M = M; a.N = N; a.K = K;
        a.ldo = epi == EPI_SWIGLU ? M / 2 : M;
        a.m_tiles = cdiv(M, BM); a.k_blocks = K / BK;
        a.stages = bn == 16 ? Smem<16>::max_stages() : bn == 32 ? Smem<32>::max_stages() : Smem<128>::max_stages();
        a.epi_full = epi; a.epi_partial = split ? EPI_ATOMIC : -1; a.hilo = hilo;
        a.lo_rows = (hilo && epi != EPI_STORE) ? bn / 2 : 0;
        const int n_tiles = hilo ? 1 : cdiv(N, bn);
        if (bn == 16) launch<16>(ta, tb, a, ctas, n_tiles, (cudaStream_t)stream);
        else if (bn == 32) launch<32>(ta, tb, a, ctas, n_tiles, (cudaStream_t)stream);
        else launch<128>(ta, tb, a, ctas, n_tiles, (cudaStream_t)stream);
        B2A_CUDA(cudaGetLastError());
        B2A_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
    });
}
