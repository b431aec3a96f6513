// probe_tld4.cu — how fast can a warp gather 2x2 fp32 footprints at 32 unrelated places? (SSAO's depth taps)
//   a) one 16-byte LDG from a precomputed quad image    b) tld4 (texture gather) from a CUDA array
//   c) tld4 from a pitch-linear 2D texture (if the driver accepts it)    d) four point fetches from a pitch-linear texture
// Also prints which texel each gather component holds. Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o probe_tld4 probe_tld4.cu
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s:%d %s -> %s\n", __FILE__, __LINE__, #x, cudaGetErrorString(e)); exit(1); } } while (0)

__device__ __forceinline__ uint32_t xs(uint32_t &s) { s ^= s << 13; s ^= s >> 17; s ^= s << 5; return s; }

template <int MODE>
__global__ void gather_kernel(const float4 *quads, const float *depth, cudaTextureObject_t tex, int W, int H, int reach, float *out) {
    const int gx = blockIdx.x * 32 + threadIdx.x, gy = blockIdx.y * 8 + threadIdx.y;
    if (gx >= W || gy >= H) return;
    uint32_t s = (gy * 9781u + gx) * 2654435761u | 1u;
    float acc = 0.0f;
#pragma unroll 4
    for (int i = 0; i < 16; ++i) {
        int x = gx + (int)(xs(s) % (2 * reach + 1)) - reach, y = gy + (int)(xs(s) % (reach + 1)) - reach / 2;
        x = min(max(x, 0), W - 2); y = min(max(y, 0), H - 2);
        float4 q;
        if (MODE == 0) q = __ldg(quads + y * W + x);
        else if (MODE == 1 || MODE == 2) q = tex2Dgather<float4>(tex, x + 1.0f, y + 1.0f, 0);
        else if (MODE == 3) q = make_float4(tex2D<float>(tex, x + 0.5f, y + 0.5f), tex2D<float>(tex, x + 1.5f, y + 0.5f), tex2D<float>(tex, x + 0.5f, y + 1.5f), tex2D<float>(tex, x + 1.5f, y + 1.5f));
        else { const float *r = depth + y * W + x; q = make_float4(__ldg(r), __ldg(r + 1), __ldg(r + W), __ldg(r + W + 1)); }
        acc += q.x + 2.0f * q.y + 3.0f * q.z + 4.0f * q.w;
    }
    out[gy * W + gx] = acc;
}

__global__ void order_kernel(cudaTextureObject_t tex, float *o) {
    float4 q = tex2Dgather<float4>(tex, 5.0f + 1.0f, 7.0f + 1.0f, 0);
    o[0] = q.x; o[1] = q.y; o[2] = q.z; o[3] = q.w;
}

int main() {
    const int W = 1920, H = 1080;
    std::vector<float> h((size_t)W * H);
    for (int y = 0; y < H; ++y) for (int x = 0; x < W; ++x) h[(size_t)y * W + x] = (float)(y * 10000 + x);
    float *d_depth, *d_out; float4 *d_quads;
    CK(cudaMalloc(&d_depth, h.size() * 4)); CK(cudaMalloc(&d_out, h.size() * 4)); CK(cudaMalloc(&d_quads, h.size() * 16));
    CK(cudaMemcpy(d_depth, h.data(), h.size() * 4, cudaMemcpyHostToDevice));
    std::vector<float4> hq(h.size());
    for (int y = 0; y < H; ++y) for (int x = 0; x < W; ++x) {
        int x1 = (x + 1) % W, y1 = (y + 1) % H;
        hq[(size_t)y * W + x] = make_float4(h[(size_t)y * W + x], h[(size_t)y * W + x1], h[(size_t)y1 * W + x], h[(size_t)y1 * W + x1]);
    }
    CK(cudaMemcpy(d_quads, hq.data(), hq.size() * 16, cudaMemcpyHostToDevice));
    // CUDA array with the gather flag
    cudaChannelFormatDesc cf = cudaCreateChannelDesc<float>();
    cudaArray_t arr;
    CK(cudaMallocArray(&arr, &cf, W, H, cudaArrayTextureGather));
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    CK(cudaMemcpy2DToArray(arr, 0, 0, d_depth, W * 4, W * 4, H, cudaMemcpyDeviceToDevice));
    CK(cudaEventRecord(e0));
    for (int i = 0; i < 20; ++i) CK(cudaMemcpy2DToArrayAsync(arr, 0, 0, d_depth, W * 4, W * 4, H, cudaMemcpyDeviceToDevice, 0));
    CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
    printf("linear -> array copy (8.3 MB): %.1f us\n", ms / 20 * 1000);
    cudaResourceDesc rd = {}; rd.resType = cudaResourceTypeArray; rd.res.array.array = arr;
    cudaTextureDesc td = {}; td.addressMode[0] = td.addressMode[1] = cudaAddressModeClamp; td.filterMode = cudaFilterModePoint;
    td.readMode = cudaReadModeElementType; td.normalizedCoords = 0;
    cudaTextureObject_t tex_arr = 0, tex_pitch = 0;
    CK(cudaCreateTextureObject(&tex_arr, &rd, &td, nullptr));
    cudaResourceDesc rp = {}; rp.resType = cudaResourceTypePitch2D; rp.res.pitch2D.devPtr = d_depth; rp.res.pitch2D.desc = cf;
    rp.res.pitch2D.width = W; rp.res.pitch2D.height = H; rp.res.pitch2D.pitchInBytes = (size_t)W * 4;
    cudaError_t pe = cudaCreateTextureObject(&tex_pitch, &rp, &td, nullptr);
    printf("pitch2D texture object: %s\n", cudaGetErrorString(pe));
    float *d_o; CK(cudaMalloc(&d_o, 16)); float ho[4];
    order_kernel<<<1, 1>>>(tex_arr, d_o); CK(cudaMemcpy(ho, d_o, 16, cudaMemcpyDeviceToHost));
    printf("array gather at footprint (5,7): x=%.0f y=%.0f z=%.0f w=%.0f   [t00=70005 t10=70006 t01=80005 t11=80006]\n", ho[0], ho[1], ho[2], ho[3]);
    if (pe == cudaSuccess) {
        order_kernel<<<1, 1>>>(tex_pitch, d_o);
        cudaError_t ge = cudaDeviceSynchronize();
        if (ge == cudaSuccess) { CK(cudaMemcpy(ho, d_o, 16, cudaMemcpyDeviceToHost)); printf("pitch gather: x=%.0f y=%.0f z=%.0f w=%.0f\n", ho[0], ho[1], ho[2], ho[3]); }
        else { printf("pitch gather failed: %s\n", cudaGetErrorString(ge)); return 0; }
    }
    dim3 block(32, 8), grid((W + 31) / 32, (H + 7) / 8);
    for (int reach : {4, 32, 144}) {
        for (int mode = 0; mode < 5; ++mode) {
            if ((mode == 2 || mode == 3) && pe != cudaSuccess) continue;
            auto run = [&]() {
                switch (mode) {
                    case 0: gather_kernel<0><<<grid, block>>>(d_quads, d_depth, tex_arr, W, H, reach, d_out); break;
                    case 1: gather_kernel<1><<<grid, block>>>(d_quads, d_depth, tex_arr, W, H, reach, d_out); break;
                    case 2: gather_kernel<2><<<grid, block>>>(d_quads, d_depth, tex_pitch, W, H, reach, d_out); break;
                    case 3: gather_kernel<3><<<grid, block>>>(d_quads, d_depth, tex_pitch, W, H, reach, d_out); break;
                    case 4: gather_kernel<4><<<grid, block>>>(d_quads, d_depth, tex_arr, W, H, reach, d_out); break;
                }
            };
            for (int i = 0; i < 3; ++i) run();
            CK(cudaEventRecord(e0));
            for (int i = 0; i < 10; ++i) run();
            CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1)); CK(cudaGetLastError());
            CK(cudaEventElapsedTime(&ms, e0, e1));
            std::vector<float> ho2(16); CK(cudaMemcpy(ho2.data(), d_out + 500 * W + 700, 64, cudaMemcpyDeviceToHost));
            static const char *names[] = {"LDG.128 quad image", "tld4 CUDA array", "tld4 pitch-linear", "4 x tex2D point pitch-linear", "4 x LDG.32 plain"};
            printf("reach +-%3d px  %-30s %7.1f us   (check %.0f)\n", reach, names[mode], ms / 10 * 1000, ho2[3]);
        }
    }
    return 0;
}
