// Issue-rate microbenchmark for the instructions of the attention kernel's softmax loop (sm_100a): which pipe each one
// occupies and whether the fp32->bf16x2 pack shares the MUFU pipe. One CTA per SM, W warps per SMSP, each thread runs
// ILP independent chains of the instruction under test for ITER iterations; reports cycles per warp-instruction per
// SMSP (reciprocal throughput) from clock64 on SM 0's CTA.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/microbench/pipes tools/microbench/pipes.cu
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

constexpr int ITER = 2048;
constexpr int ILP = 8;

enum Op { EX2, CVT, EX2_CVT, FMNMX, FMNMX3, FFMA2, FADD2, FFMA, FADDRM, PRMT, SHL, IADD, EX2_PRMT, LOP, EX2_BF16X2, EX2_F16X2,
          HFMA2_BF16, NOPS };
static const char* names[] = {"ex2.approx", "cvt.bf16x2", "ex2+cvt(1:0.5)", "max.f32", "max.f32 x3", "fma.f32x2",
                              "add.f32x2", "fma.f32", "add.rm.f32", "prmt", "shl", "add.s32", "ex2+prmt(1:0.5)",
                              "and.b32", "ex2.bf16x2", "ex2.f16x2", "fma.bf16x2"};

template <int OP>
__global__ void __launch_bounds__(1024) k(float* out, long long* cyc, float seed) {
  float a[ILP], b[ILP];
  uint32_t u[ILP];
#pragma unroll
  for (int i = 0; i < ILP; ++i) {
    a[i] = seed + threadIdx.x * 1e-3f + i;
    b[i] = seed * 0.5f + i;
    u[i] = threadIdx.x + i;
  }
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < ITER; ++it) {
#pragma unroll
    for (int i = 0; i < ILP; ++i) {
      if (OP == EX2) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[i]));
      if (OP == CVT) asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(u[i]) : "f"(a[i]), "f"(b[i]));
      if (OP == EX2_CVT) {
        asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[i]));
        if (i & 1) asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(u[i]) : "f"(a[i]), "f"(a[i - 1]));
      }
      if (OP == EX2_PRMT) {
        asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[i]));
        if (i & 1)
          asm volatile("prmt.b32 %0, %1, %2, 0x7632;" : "=r"(u[i]) : "r"(__float_as_uint(a[i - 1])), "r"(__float_as_uint(a[i])));
      }
      if (OP == EX2_BF16X2) asm volatile("ex2.approx.ftz.bf16x2 %0, %0;" : "+r"(u[i]));
      if (OP == EX2_F16X2) asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(u[i]));
      if (OP == HFMA2_BF16) asm volatile("fma.rn.bf16x2 %0, %0, %1, %2;" : "+r"(u[i]) : "r"(u[(i + 1) % ILP]), "r"(u[(i + 2) % ILP]));
      if (OP == FMNMX) asm volatile("max.f32 %0, %0, %1;" : "+f"(a[i]) : "f"(b[i]));
      if (OP == FMNMX3) asm volatile("max.f32 %0, %0, %1, %2;" : "+f"(a[i]) : "f"(b[i]), "f"(b[(i + 1) % ILP]));
      if (OP == FFMA2) {
        if (!(i & 1)) {
          uint64_t d, x, y, z;
          asm volatile("mov.b64 %0, {%1, %2};" : "=l"(x) : "f"(a[i]), "f"(a[i + 1]));
          asm volatile("mov.b64 %0, {%1, %2};" : "=l"(y) : "f"(b[i]), "f"(b[i + 1]));
          asm volatile("mov.b64 %0, {%1, %2};" : "=l"(z) : "f"(b[i + 1]), "f"(b[i]));
          asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(x), "l"(y), "l"(z));
          asm volatile("mov.b64 {%0, %1}, %2;" : "=f"(a[i]), "=f"(a[i + 1]) : "l"(d));
        }
      }
      if (OP == FADD2) {
        if (!(i & 1)) {
          uint64_t d, x, y;
          asm volatile("mov.b64 %0, {%1, %2};" : "=l"(x) : "f"(a[i]), "f"(a[i + 1]));
          asm volatile("mov.b64 %0, {%1, %2};" : "=l"(y) : "f"(b[i]), "f"(b[i + 1]));
          asm volatile("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(x), "l"(y));
          asm volatile("mov.b64 {%0, %1}, %2;" : "=f"(a[i]), "=f"(a[i + 1]) : "l"(d));
        }
      }
      if (OP == FFMA) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(a[i]) : "f"(b[i]), "f"(b[(i + 1) % ILP]));
      if (OP == FADDRM) asm volatile("add.rm.ftz.f32 %0, %0, %1;" : "+f"(a[i]) : "f"(b[i]));
      if (OP == PRMT) asm volatile("prmt.b32 %0, %0, %1, 0x7632;" : "+r"(u[i]) : "r"(u[(i + 1) % ILP]));
      if (OP == SHL) asm volatile("shl.b32 %0, %0, 1;" : "+r"(u[i]));
      if (OP == IADD) asm volatile("add.s32 %0, %0, %1;" : "+r"(u[i]) : "r"(u[(i + 1) % ILP]));
      if (OP == LOP) asm volatile("and.b32 %0, %0, %1;" : "+r"(u[i]) : "r"(u[(i + 1) % ILP] | 0xffff0000u));
    }
  }
  const long long t1 = clock64();
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < ILP; ++i) s += a[i] + __uint_as_float(u[i]);
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (blockIdx.x == 0 && threadIdx.x == 0) *cyc = t1 - t0;
}

template <int OP>
void run(int warps_per_smsp, float* out, long long* cyc) {
  const int threads = warps_per_smsp * 4 * 32;
  k<OP><<<148, threads>>>(out, cyc, 1.0f);
  cudaDeviceSynchronize();
  k<OP><<<148, threads>>>(out, cyc, 1.0f);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) {
    printf("%-18s error %s\n", names[OP], cudaGetErrorString(e));
    return;
  }
  long long c;
  cudaMemcpy(&c, cyc, sizeof(c), cudaMemcpyDeviceToHost);
  double per_iter_instr = ILP;  // "instruction groups" per iteration per warp
  if (OP == FFMA2 || OP == FADD2) per_iter_instr = ILP / 2;
  printf("%-18s warps/SMSP=%d  %.2f cyc per warp-instr(-group) per SMSP\n", names[OP], warps_per_smsp,
         (double)c / ((double)ITER * per_iter_instr * warps_per_smsp));
}

int main() {
  float* out;
  long long* cyc;
  cudaMalloc(&out, 148 * 1024 * sizeof(float));
  cudaMalloc(&cyc, sizeof(long long));
  for (int w : {2, 4}) {
    run<EX2>(w, out, cyc);
    run<CVT>(w, out, cyc);
    run<EX2_CVT>(w, out, cyc);
    run<EX2_PRMT>(w, out, cyc);
    run<FMNMX>(w, out, cyc);
    run<FMNMX3>(w, out, cyc);
    run<FFMA2>(w, out, cyc);
    run<FADD2>(w, out, cyc);
    run<FFMA>(w, out, cyc);
    run<FADDRM>(w, out, cyc);
    run<PRMT>(w, out, cyc);
    run<SHL>(w, out, cyc);
    run<IADD>(w, out, cyc);
    run<LOP>(w, out, cyc);
    run<EX2_BF16X2>(w, out, cyc);
    run<EX2_F16X2>(w, out, cyc);
    run<HFMA2_BF16>(w, out, cyc);
  }
  return 0;
}
