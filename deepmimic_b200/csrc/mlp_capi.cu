// C ABI of the tensor-core policy network (include/deepmimic_b200.h, dm_mlp_*): host-side weight tiling + the four launches (operand preparation, three GEMMs) of
// kernels/dm_mlp.cu.  Same library, same rules: no CPU fallback, errors through dm_last_error.
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include <cmath>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/deepmimic_b200.h"

namespace dmk {
struct MlpPrepParams { const float* obs; const float* in_mean; const float* in_istd; float in_clip; int in_dim, M, NC; __half* tiles; };
struct MlpGemmParams {
    const __half* a_tiles; const __half* w_tiles; const float* bias; __half* out_tiles; float* actions; const float* out_mean; const float* out_std; const float* noise;
    int out_dim; int M, K, N;
};
__global__ void dm_mlp_prep_kernel(MlpPrepParams);
template <int BN, bool LAST>
__global__ void dm_mlp_gemm_kernel(MlpGemmParams);
int dm_mlp_smem_bytes(int bn);
}  // namespace dmk

extern "C" void dm_set_last_error(const char* msg);

struct dm_mlp {
    int device = 0, in_dim = 0, h0 = 0, h1 = 0, out_dim = 0, max_rows = 0;
    int K0 = 0, N0 = 0, N1 = 0, N2 = 0;   // padded sizes: K0 = pad64(in), N0 = pad128(h0) = K1, N1 = pad128(h1) = K2, N2 = 32 / 64
    __half *w[3] = {nullptr, nullptr, nullptr}, *obs_t = nullptr, *act0 = nullptr, *act1 = nullptr;   // activations: operand tiles [m tiles][K / 64][128 x 64]
    float *b[3] = {nullptr, nullptr, nullptr}, *in_mean = nullptr, *in_istd = nullptr, *out_mean = nullptr, *out_std = nullptr;
    float in_clip = 1e30f;
    long long launches = 0;
};

namespace {
int mlp_fail(const std::string& m) { dm_set_last_error(m.c_str()); std::fprintf(stderr, "[deepmimic_b200] %s\n", m.c_str()); return 1; }
int pad_to(int v, int q) { return ((v + q - 1) / q) * q; }
// w: [K_in x N_out] row major (the reference's dense kernels: inputs x units).  Tiles: [n tile][k chunk][hi | lo][k8][row group][row][8 halves]
std::vector<__half> tile_weights(const float* w, int k_in, int n_out, int K, int N, int BN) {
    const int NC = K / 64, NT = N / BN;
    std::vector<__half> out(static_cast<size_t>(NT) * NC * 2 * BN * 64);
    for (int nt = 0; nt < NT; ++nt)
        for (int c = 0; c < NC; ++c) {
            __half* hi = &out[(static_cast<size_t>(nt) * NC + c) * 2 * BN * 64];
            __half* lo = hi + BN * 64;
            for (int k8 = 0; k8 < 8; ++k8)
                for (int rg = 0; rg < BN / 8; ++rg)
                    for (int r = 0; r < 8; ++r)
                        for (int e = 0; e < 8; ++e) {
                            const int n = nt * BN + rg * 8 + r, k = c * 64 + k8 * 8 + e;
                            const float v = (n < n_out && k < k_in) ? w[static_cast<size_t>(k) * n_out + n] : 0.f;
                            const __half h = __float2half_rn(v);
                            const size_t o = (static_cast<size_t>(k8) * (BN / 8) + rg) * 64 + r * 8 + e;
                            hi[o] = h; lo[o] = __float2half_rn(v - __half2float(h));
                        }
        }
    return out;
}
template <class T>
bool upload(T** dst, const std::vector<T>& src) {
    if (cudaMalloc(dst, src.size() * sizeof(T)) != cudaSuccess) return false;
    return cudaMemcpy(*dst, src.data(), src.size() * sizeof(T), cudaMemcpyHostToDevice) == cudaSuccess;
}
std::vector<float> padded(const float* v, int n, int N, float fill = 0.f) { std::vector<float> o(N, fill); if (v) std::memcpy(o.data(), v, sizeof(float) * n); return o; }
}  // namespace

extern "C" {

dm_mlp* dm_mlp_create(int device, int in_dim, int h0, int h1, int out_dim, const float* w0, const float* b0, const float* w1, const float* b1, const float* w2, const float* b2,
                      const float* in_mean, const float* in_std, float in_clip, const float* out_mean, const float* out_std, int max_rows) {
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) { mlp_fail("dm_mlp_create: no CUDA device (the policy network has no CPU fallback)"); return nullptr; }
    if (in_dim <= 0 || h0 <= 0 || h1 <= 0 || out_dim <= 0 || out_dim > 64 || max_rows <= 0) { mlp_fail("dm_mlp_create: bad sizes (out_dim must be <= 64)"); return nullptr; }
    if (cudaSetDevice(device) != cudaSuccess) { mlp_fail("dm_mlp_create: cudaSetDevice failed"); return nullptr; }
    cudaDeviceProp prop;
    cudaGetDeviceProperties(&prop, device);
    if (prop.major < 10) { mlp_fail("dm_mlp_create: tcgen05 tensor cores need sm_100a (found sm_" + std::to_string(prop.major) + std::to_string(prop.minor) + ")"); return nullptr; }
    dm_mlp* m = new dm_mlp();
    m->device = device; m->in_dim = in_dim; m->h0 = h0; m->h1 = h1; m->out_dim = out_dim; m->max_rows = pad_to(max_rows, 128);
    m->K0 = pad_to(in_dim, 64); m->N0 = pad_to(h0, 128); m->N1 = pad_to(h1, 128); m->N2 = out_dim <= 32 ? 32 : 64;
    m->in_clip = in_clip > 0.f ? in_clip : 1e30f;
    std::vector<float> istd(in_dim, 1.f);
    for (int i = 0; i < in_dim; ++i) istd[i] = in_std ? 1.0f / in_std[i] : 1.f;
    bool ok = upload(&m->w[0], tile_weights(w0, in_dim, h0, m->K0, m->N0, 128)) && upload(&m->w[1], tile_weights(w1, h0, h1, m->N0, m->N1, 128)) &&
              upload(&m->w[2], tile_weights(w2, h1, out_dim, m->N1, m->N2, m->N2)) && upload(&m->b[0], padded(b0, h0, m->N0)) && upload(&m->b[1], padded(b1, h1, m->N1)) &&
              upload(&m->b[2], padded(b2, out_dim, m->N2)) && upload(&m->in_mean, padded(in_mean, in_dim, in_dim)) && upload(&m->in_istd, istd) &&
              upload(&m->out_mean, padded(out_mean, out_dim, out_dim)) && upload(&m->out_std, padded(out_std, out_dim, out_dim, 1.f)) &&
              cudaMalloc(&m->obs_t, static_cast<size_t>(m->max_rows) * m->K0 * sizeof(__half)) == cudaSuccess &&
              cudaMalloc(&m->act0, static_cast<size_t>(m->max_rows) * m->N0 * sizeof(__half)) == cudaSuccess &&
              cudaMalloc(&m->act1, static_cast<size_t>(m->max_rows) * m->N1 * sizeof(__half)) == cudaSuccess;
    if (ok && !out_std) { std::vector<float> one(out_dim, 1.f); ok = cudaMemcpy(m->out_std, one.data(), sizeof(float) * out_dim, cudaMemcpyHostToDevice) == cudaSuccess; }
    if (ok) {
        ok = cudaFuncSetAttribute(dmk::dm_mlp_gemm_kernel<128, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, dmk::dm_mlp_smem_bytes(128)) == cudaSuccess &&
             cudaFuncSetAttribute(dmk::dm_mlp_gemm_kernel<32, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, dmk::dm_mlp_smem_bytes(32)) == cudaSuccess &&
             cudaFuncSetAttribute(dmk::dm_mlp_gemm_kernel<64, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, dmk::dm_mlp_smem_bytes(64)) == cudaSuccess;
    }
    if (!ok) { mlp_fail(std::string("dm_mlp_create: ") + cudaGetErrorString(cudaGetLastError())); dm_mlp_destroy(m); return nullptr; }
    return m;
}

int dm_mlp_forward(dm_mlp* m, const float* d_obs, const float* d_noise, float* d_actions, int rows, void* stream) {
    if (!m) return mlp_fail("dm_mlp_forward: null handle");
    if (rows <= 0 || rows > m->max_rows) return mlp_fail("dm_mlp_forward: rows out of range");
    if (cudaSetDevice(m->device) != cudaSuccess) return mlp_fail("dm_mlp_forward: cudaSetDevice failed");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int mt = (rows + 127) / 128;
    // observations -> normalised fp16 operand tiles
    dmk::MlpPrepParams Q{d_obs, m->in_mean, m->in_istd, m->in_clip, m->in_dim, rows, m->K0 / 64, m->obs_t};
    dmk::dm_mlp_prep_kernel<<<dim3(mt, m->K0 / 64), 128, 0, st>>>(Q);
    dmk::MlpGemmParams P{};
    P.M = rows;
    // layer 0: 227 -> 1024 + ReLU
    P.a_tiles = m->obs_t; P.w_tiles = m->w[0]; P.bias = m->b[0]; P.out_tiles = m->act0; P.K = m->K0; P.N = m->N0;
    dmk::dm_mlp_gemm_kernel<128, false><<<dim3(mt, m->N0 / 128), 128, dmk::dm_mlp_smem_bytes(128), st>>>(P);
    // layer 1: 1024 -> 512 + ReLU
    P.a_tiles = m->act0; P.w_tiles = m->w[1]; P.bias = m->b[1]; P.out_tiles = m->act1; P.K = m->N0; P.N = m->N1;
    dmk::dm_mlp_gemm_kernel<128, false><<<dim3(mt, m->N1 / 128), 128, dmk::dm_mlp_smem_bytes(128), st>>>(P);
    // layer 2: 512 -> actions, un-normalised
    P.a_tiles = m->act1; P.w_tiles = m->w[2]; P.bias = m->b[2]; P.out_tiles = nullptr; P.actions = d_actions; P.out_mean = m->out_mean; P.out_std = m->out_std; P.noise = d_noise;
    P.out_dim = m->out_dim; P.K = m->N1; P.N = m->N2;
    if (m->N2 == 32) dmk::dm_mlp_gemm_kernel<32, true><<<dim3(mt, 1), 128, dmk::dm_mlp_smem_bytes(32), st>>>(P);
    else dmk::dm_mlp_gemm_kernel<64, true><<<dim3(mt, 1), 128, dmk::dm_mlp_smem_bytes(64), st>>>(P);
    const cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return mlp_fail(std::string("dm_mlp_forward: ") + cudaGetErrorString(e));
    m->launches += 4;
    return 0;
}

long long dm_mlp_launches(dm_mlp* m) { return m ? m->launches : 0; }

void dm_mlp_destroy(dm_mlp* m) {
    if (!m) return;
    cudaSetDevice(m->device);
    for (auto& p : m->w) cudaFree(p);
    for (auto& p : m->b) cudaFree(p);
    cudaFree(m->obs_t); cudaFree(m->act0); cudaFree(m->act1); cudaFree(m->in_mean); cudaFree(m->in_istd); cudaFree(m->out_mean); cudaFree(m->out_std);
    delete m;
}

}  // extern "C"
