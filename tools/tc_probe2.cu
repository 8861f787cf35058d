// tc_probe2.cu -- bring-up probe: tcgen05.mma kind::tf32 with BOTH operands from shared memory in the MN-major
// canonical layout (the dW = A^T dP product: contraction over observations), M = 64, N = 64, K = 256.
// Dumps all 128 TMEM lanes so the host can discover where the 64 accumulator rows live.
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cmath>
#include <vector>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
constexpr int KOBS = 256, MF = 64, NF = 64;
constexpr uint32_t LBO = 128;                 // between 8-row K groups
constexpr uint32_t SBO = (KOBS / 8) * 128;    // between 4-element MN groups
__host__ __device__ inline uint32_t mn_off(int mn, int k) { return (mn % 4) * 4 + (k % 8) * 16 + (mn / 4) * SBO + (k / 8) * LBO; }
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
  return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)((lbo >> 4) & 0x3FFF) << 16) | ((uint64_t)((sbo >> 4) & 0x3FFF) << 32) | ((uint64_t)1 << 46);
}
__device__ __forceinline__ void mma_ss(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n" :: "r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile("{\n\t.reg .pred P1;\n\tWL:\n\tmbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t@P1 bra DN;\n\tbra WL;\n\tDN:\n\t}\n" :: "r"(smem_u32(bar)), "r"(parity) : "memory");
}
#define TMEM_LD32(taddr, v) asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];" \
  : "=r"(v[0]),"=r"(v[1]),"=r"(v[2]),"=r"(v[3]),"=r"(v[4]),"=r"(v[5]),"=r"(v[6]),"=r"(v[7]),"=r"(v[8]),"=r"(v[9]),"=r"(v[10]),"=r"(v[11]),"=r"(v[12]),"=r"(v[13]),"=r"(v[14]),"=r"(v[15]), \
    "=r"(v[16]),"=r"(v[17]),"=r"(v[18]),"=r"(v[19]),"=r"(v[20]),"=r"(v[21]),"=r"(v[22]),"=r"(v[23]),"=r"(v[24]),"=r"(v[25]),"=r"(v[26]),"=r"(v[27]),"=r"(v[28]),"=r"(v[29]),"=r"(v[30]),"=r"(v[31]) \
  : "r"(taddr) : "memory")

constexpr uint32_t K_SBO = 128, K_LBO = (64 / 8) * 128;    // K-major image of a [64][256] operand
__host__ __device__ inline uint32_t k_off(int mn, int k) { return (k / 4) * K_LBO + (mn / 8) * K_SBO + (mn % 8) * 16 + (k % 4) * 4; }
// MN-major, 128-byte swizzle: atoms of 8 k-rows x 128 B (32 MN elements), 16-byte chunk index XOR (row % 8)
constexpr uint32_t W_SBO = 1024, W_LBO = (KOBS / 8) * 1024;
__host__ __device__ inline uint32_t sw_off(int mn, int k) {
  const int chunk = (mn % 32) / 4, row = k % 8;
  return (mn / 32) * W_LBO + (k / 8) * W_SBO + row * 128 + ((chunk ^ row) * 16) + (mn % 4) * 4;
}
// MN-major tf32: SWIZZLE_128B_BASE32B (layout type 1): atoms of 4 k-rows x 128 B (32 MN elements), 32-byte chunk index XOR (row % 4)
constexpr uint32_t B_SBO = 512, B_LBO = (KOBS / 4) * 512;
__host__ __device__ inline uint32_t b32_off(int mn, int k) {
  const int c = (mn % 32) / 8, r = k % 4;
  return (mn / 32) * B_LBO + (k / 4) * B_SBO + r * 128 + ((c ^ r) * 32) + (mn % 8) * 4;
}
__device__ __forceinline__ uint64_t make_desc_b32(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
  return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)((lbo >> 4) & 0x3FFF) << 16) | ((uint64_t)((sbo >> 4) & 0x3FFF) << 32) | ((uint64_t)1 << 46) | ((uint64_t)1 << 61);
}
__device__ __forceinline__ uint64_t make_desc_sw128(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
  return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)((lbo >> 4) & 0x3FFF) << 16) | ((uint64_t)((sbo >> 4) & 0x3FFF) << 32) | ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}
__global__ void __launch_bounds__(128, 1) probe(const float* X, const float* Y, float* Dout, int M, int amaj, int bmaj) {
  extern __shared__ __align__(1024) unsigned char smem[];
  float* Aimg = reinterpret_cast<float*>(smem);                    // 64 KB
  float* Bimg = reinterpret_cast<float*>(smem + 65536);            // 64 KB
  __shared__ __align__(8) uint64_t mbar;
  __shared__ uint32_t tbase_s;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int idx = tid; idx < KOBS * MF; idx += 128) {
    const int k = idx / MF, m = idx % MF;
    *reinterpret_cast<float*>(reinterpret_cast<char*>(Aimg) + (amaj == 3 ? b32_off(m, k) : amaj == 2 ? sw_off(m, k) : amaj ? mn_off(m, k) : k_off(m, k))) = X[k * MF + m];
    *reinterpret_cast<float*>(reinterpret_cast<char*>(Bimg) + (bmaj == 3 ? b32_off(m, k) : bmaj == 2 ? sw_off(m, k) : bmaj ? mn_off(m, k) : k_off(m, k))) = Y[k * NF + m];
  }
  if (tid == 0) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(smem_u32(&mbar)) : "memory");
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 128;" :: "r"(smem_u32(&tbase_s)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tb = tbase_s;
  // zero the accumulator columns of all 128 lanes first (so untouched lanes read back as 0)
  {
    uint32_t z[32];
    for (int i = 0; i < 32; ++i) z[i] = 0;
    const uint32_t row = tb + ((uint32_t)(32 * warp) << 16);
    asm volatile("tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32};"
      :: "r"(row), "r"(z[0]),"r"(z[1]),"r"(z[2]),"r"(z[3]),"r"(z[4]),"r"(z[5]),"r"(z[6]),"r"(z[7]),"r"(z[8]),"r"(z[9]),"r"(z[10]),"r"(z[11]),"r"(z[12]),"r"(z[13]),"r"(z[14]),"r"(z[15]),
         "r"(z[16]),"r"(z[17]),"r"(z[18]),"r"(z[19]),"r"(z[20]),"r"(z[21]),"r"(z[22]),"r"(z[23]),"r"(z[24]),"r"(z[25]),"r"(z[26]),"r"(z[27]),"r"(z[28]),"r"(z[29]),"r"(z[30]),"r"(z[31]) : "memory");
    asm volatile("tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32};"
      :: "r"(row + 32), "r"(z[0]),"r"(z[1]),"r"(z[2]),"r"(z[3]),"r"(z[4]),"r"(z[5]),"r"(z[6]),"r"(z[7]),"r"(z[8]),"r"(z[9]),"r"(z[10]),"r"(z[11]),"r"(z[12]),"r"(z[13]),"r"(z[14]),"r"(z[15]),
         "r"(z[16]),"r"(z[17]),"r"(z[18]),"r"(z[19]),"r"(z[20]),"r"(z[21]),"r"(z[22]),"r"(z[23]),"r"(z[24]),"r"(z[25]),"r"(z[26]),"r"(z[27]),"r"(z[28]),"r"(z[29]),"r"(z[30]),"r"(z[31]) : "memory");
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (tid == 0) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    // idesc: D=F32, A=B=TF32, A and B MN-major (bits 15, 16), N=64, M
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(amaj != 0) << 15) | ((uint32_t)(bmaj != 0) << 16) | ((uint32_t)(NF >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
    const uint64_t da = amaj == 3 ? make_desc_b32(smem_u32(Aimg), B_LBO, B_SBO) : amaj == 2 ? make_desc_sw128(smem_u32(Aimg), W_LBO, W_SBO) : amaj ? make_desc(smem_u32(Aimg), LBO, SBO) : make_desc(smem_u32(Aimg), K_LBO, K_SBO);
    const uint64_t db = bmaj == 3 ? make_desc_b32(smem_u32(Bimg), B_LBO, B_SBO) : bmaj == 2 ? make_desc_sw128(smem_u32(Bimg), W_LBO, W_SBO) : bmaj ? make_desc(smem_u32(Bimg), LBO, SBO) : make_desc(smem_u32(Bimg), K_LBO, K_SBO);
    const uint32_t sa = amaj == 3 ? 2 * B_SBO : amaj == 2 ? W_SBO : amaj ? LBO : 2 * K_LBO, sb = bmaj == 3 ? 2 * B_SBO : bmaj == 2 ? W_SBO : bmaj ? LBO : 2 * K_LBO;
    for (int ks = 0; ks < KOBS / 8; ++ks)
      mma_ss(tb, da + (uint64_t)((sa * ks) >> 4), db + (uint64_t)((sb * ks) >> 4), idesc, ks > 0 ? 1u : 0u);
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(smem_u32(&mbar)) : "memory");
  }
  mbar_wait(&mbar, 0);
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  uint32_t v[32];
  const uint32_t row = tb + ((uint32_t)(32 * warp) << 16);
  TMEM_LD32(row, v);
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
  for (int c = 0; c < 32; ++c) Dout[tid * 64 + c] = __uint_as_float(v[c]);
  TMEM_LD32(row + 32, v);
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
  for (int c = 0; c < 32; ++c) Dout[tid * 64 + 32 + c] = __uint_as_float(v[c]);
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 128;" :: "r"(tb) : "memory");
}

int main() {
  std::vector<float> X(KOBS * MF), Y(KOBS * NF), D(128 * 64);
  srand(3);
  for (auto& x : X) x = (float)((rand() % 17) - 8) / 8.0f;       // exactly representable in tf32
  for (auto& y : Y) y = (float)((rand() % 17) - 8) / 8.0f;
  std::vector<double> ref(MF * NF, 0.0);
  for (int k = 0; k < KOBS; ++k) for (int m = 0; m < MF; ++m) for (int n = 0; n < NF; ++n) ref[m * NF + n] += (double)X[k * MF + m] * Y[k * NF + n];
  float *dX, *dY, *dD;
  CK(cudaMalloc(&dX, X.size() * 4)); CK(cudaMalloc(&dY, Y.size() * 4)); CK(cudaMalloc(&dD, D.size() * 4));
  CK(cudaMemcpy(dX, X.data(), X.size() * 4, cudaMemcpyHostToDevice)); CK(cudaMemcpy(dY, Y.data(), Y.size() * 4, cudaMemcpyHostToDevice));
  CK(cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 132 * 1024));
  const int combos[][3] = {{64, 0, 0}, {64, 3, 0}, {64, 0, 3}, {64, 3, 3}};
  for (auto& cb : combos) {
    const int M = cb[0], amaj = cb[1], bmaj = cb[2];
    probe<<<1, 128, 132 * 1024>>>(dX, dY, dD, M, amaj, bmaj);
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost));
    printf("M=%d A %s-major, B %s-major: lane -> matching reference row (max |err| over 64 cols):\n", M, amaj == 3 ? "MN-sw128base32" : amaj == 2 ? "MN-sw128" : amaj ? "MN" : "K", bmaj == 3 ? "MN-sw128base32" : bmaj == 2 ? "MN-sw128" : bmaj ? "MN" : "K");
    int found = 0;
    for (int lane = 0; lane < 128; ++lane) {
      int best = -1; double beste = 1e30; bool nz = false;
      for (int c = 0; c < 64; ++c) nz |= (D[lane * 64 + c] != 0.f);
      for (int m = 0; m < MF; ++m) {
        double e = 0; for (int c = 0; c < 64; ++c) e = fmax(e, fabs(D[lane * 64 + c] - ref[m * NF + c]));
        if (e < beste) { beste = e; best = m; }
      }
      if (nz) { if (found < 8 || lane % 16 == 0) printf(" [lane %3d -> row %2d err %.1e]%s", lane, best, beste, (found % 4 == 3) ? "\n" : ""); ++found; }
    }
    printf("\n  non-zero lanes: %d\n", found);
  }
  return 0;
}
