// Issue-rate microbenchmark for the instruction mix of the Gotoh cell update on sm_100a.
// One CTA of 512 threads (16 warps, 4 per SMSP) per SM; every thread runs ITER iterations of
// 8 independent chains of the op under test; cycles from clock64() inside the CTA (max over CTAs).
// Prints warp-instructions / cycle / SM (issue peak = 4.0) and lane-ops / cycle / SM.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define ITER 4096

template <int OP>
__global__ void __launch_bounds__(512) bench(float* out, long long* cycles, float c0, float c1) {
    float x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
    unsigned long long p0 = threadIdx.x, p1 = p0 + 1, p2 = p0 + 2, p3 = p0 + 3, cc;
    unsigned u0 = threadIdx.x, u1 = u0 + 1, u2 = u0 + 2, u3 = u0 + 3;
    {
        float2 c = make_float2(c0, c1);
        cc = *reinterpret_cast<unsigned long long*>(&c);
    }
    __shared__ float4 sm[512];
    sm[threadIdx.x] = make_float4(c0, c1, c0, c1);
    __syncthreads();
    long long t0 = clock64();
#pragma unroll 1
    for(int it = 0; it < ITER; ++it) {
        if(OP == 0) {  // FADD x8
            asm volatile("add.rn.f32 %0, %0, %8; add.rn.f32 %1, %1, %8; add.rn.f32 %2, %2, %8; add.rn.f32 %3, %3, %8;"
                         "add.rn.f32 %4, %4, %8; add.rn.f32 %5, %5, %8; add.rn.f32 %6, %6, %8; add.rn.f32 %7, %7, %8;"
                         : "+f"(x0), "+f"(x1), "+f"(x2), "+f"(x3), "+f"(x4), "+f"(x5), "+f"(x6), "+f"(x7) : "f"(c0));
        } else if(OP == 1) {  // FADD2 x4 (8 lane-adds) twice => 8 instr
            asm volatile("add.rn.f32x2 %0, %0, %4; add.rn.f32x2 %1, %1, %4; add.rn.f32x2 %2, %2, %4; add.rn.f32x2 %3, %3, %4;"
                         "add.rn.f32x2 %0, %0, %4; add.rn.f32x2 %1, %1, %4; add.rn.f32x2 %2, %2, %4; add.rn.f32x2 %3, %3, %4;"
                         : "+l"(p0), "+l"(p1), "+l"(p2), "+l"(p3) : "l"(cc));
        } else if(OP == 2) {  // FMNMX x8
            asm volatile("max.f32 %0, %0, %8; max.f32 %1, %1, %8; max.f32 %2, %2, %8; max.f32 %3, %3, %8;"
                         "max.f32 %4, %4, %8; max.f32 %5, %5, %8; max.f32 %6, %6, %8; max.f32 %7, %7, %8;"
                         : "+f"(x0), "+f"(x1), "+f"(x2), "+f"(x3), "+f"(x4), "+f"(x5), "+f"(x6), "+f"(x7) : "f"(c0));
        } else if(OP == 3) {  // FMNMX3 x8
            asm volatile("max.f32 %0, %0, %8, %9; max.f32 %1, %1, %8, %9; max.f32 %2, %2, %8, %9; max.f32 %3, %3, %8, %9;"
                         "max.f32 %4, %4, %8, %9; max.f32 %5, %5, %8, %9; max.f32 %6, %6, %8, %9; max.f32 %7, %7, %8, %9;"
                         : "+f"(x0), "+f"(x1), "+f"(x2), "+f"(x3), "+f"(x4), "+f"(x5), "+f"(x6), "+f"(x7) : "f"(c0), "f"(c1));
        } else if(OP == 4) {  // SHF x8
            asm volatile("shf.l.wrap.b32 %0, %4, %0, 1; shf.l.wrap.b32 %1, %4, %1, 1; shf.l.wrap.b32 %2, %4, %2, 1; shf.l.wrap.b32 %3, %4, %3, 1;"
                         "shf.l.wrap.b32 %0, %5, %0, 1; shf.l.wrap.b32 %1, %5, %1, 1; shf.l.wrap.b32 %2, %5, %2, 1; shf.l.wrap.b32 %3, %5, %3, 1;"
                         : "+r"(u0), "+r"(u1), "+r"(u2), "+r"(u3) : "r"(__float_as_uint(c0)), "r"(__float_as_uint(c1)));
        } else if(OP == 5) {  // 4 FADD + 4 FMNMX interleaved
            asm volatile("add.rn.f32 %0, %0, %8; max.f32 %1, %1, %8; add.rn.f32 %2, %2, %8; max.f32 %3, %3, %8;"
                         "add.rn.f32 %4, %4, %8; max.f32 %5, %5, %8; add.rn.f32 %6, %6, %8; max.f32 %7, %7, %8;"
                         : "+f"(x0), "+f"(x1), "+f"(x2), "+f"(x3), "+f"(x4), "+f"(x5), "+f"(x6), "+f"(x7) : "f"(c0));
        } else if(OP == 6) {  // 4 FADD2 + 4 FMNMX interleaved (12 lane-ops in 8 instr)
            asm volatile("add.rn.f32x2 %0, %0, %8; max.f32 %4, %4, %9; add.rn.f32x2 %1, %1, %8; max.f32 %5, %5, %9;"
                         "add.rn.f32x2 %2, %2, %8; max.f32 %6, %6, %9; add.rn.f32x2 %3, %3, %8; max.f32 %7, %7, %9;"
                         : "+l"(p0), "+l"(p1), "+l"(p2), "+l"(p3), "+f"(x4), "+f"(x5), "+f"(x6), "+f"(x7) : "l"(cc), "f"(c0));
        } else if(OP == 7) {  // 4 FADD + 4 SHF interleaved
            asm volatile("add.rn.f32 %0, %0, %8; shf.l.wrap.b32 %4, %9, %4, 1; add.rn.f32 %1, %1, %8; shf.l.wrap.b32 %5, %9, %5, 1;"
                         "add.rn.f32 %2, %2, %8; shf.l.wrap.b32 %6, %9, %6, 1; add.rn.f32 %3, %3, %8; shf.l.wrap.b32 %7, %9, %7, 1;"
                         : "+f"(x0), "+f"(x1), "+f"(x2), "+f"(x3), "+r"(u0), "+r"(u1), "+r"(u2), "+r"(u3) : "f"(c0), "r"(__float_as_uint(c1)));
        } else if(OP == 8) {  // FSETP + SELP pairs (4 of each)
            asm volatile("{ .reg .pred q0, q1, q2, q3;"
                         "setp.gt.f32 q0, %0, %8; setp.gt.f32 q1, %1, %8; setp.gt.f32 q2, %2, %8; setp.gt.f32 q3, %3, %8;"
                         "selp.f32 %4, %4, %9, q0; selp.f32 %5, %5, %9, q1; selp.f32 %6, %6, %9, q2; selp.f32 %7, %7, %9, q3; }"
                         : "+f"(x0), "+f"(x1), "+f"(x2), "+f"(x3), "+f"(x4), "+f"(x5), "+f"(x6), "+f"(x7) : "f"(c0), "f"(c1));
        } else if(OP == 9) {  // SHFL x8
            x0 = __shfl_up_sync(0xffffffffu, x0, 1); x1 = __shfl_up_sync(0xffffffffu, x1, 1);
            x2 = __shfl_up_sync(0xffffffffu, x2, 1); x3 = __shfl_up_sync(0xffffffffu, x3, 1);
            x4 = __shfl_up_sync(0xffffffffu, x4, 1); x5 = __shfl_up_sync(0xffffffffu, x5, 1);
            x6 = __shfl_up_sync(0xffffffffu, x6, 1); x7 = __shfl_up_sync(0xffffffffu, x7, 1);
        } else if(OP == 10) {  // LDS.128 x2 + 6 FADD
            float4 v = sm[(threadIdx.x + it) & 511], w = sm[(threadIdx.x + 2 * it + 7) & 511];
            x0 += v.x; x1 += v.y; x2 += v.z; x3 += v.w; x4 += w.x; x5 += w.y;
        } else if(OP == 11) {  // LOP3 x8
            asm volatile("lop3.b32 %0, %0, %4, %5, 0x96; lop3.b32 %1, %1, %4, %5, 0x96; lop3.b32 %2, %2, %4, %5, 0x96; lop3.b32 %3, %3, %4, %5, 0x96;"
                         "lop3.b32 %0, %0, %5, %4, 0xe8; lop3.b32 %1, %1, %5, %4, 0xe8; lop3.b32 %2, %2, %5, %4, 0xe8; lop3.b32 %3, %3, %5, %4, 0xe8;"
                         : "+r"(u0), "+r"(u1), "+r"(u2), "+r"(u3) : "r"(__float_as_uint(c0)), "r"(__float_as_uint(c1)));
        } else if(OP == 12) {  // realistic mix: 4 FADD2 (8 adds) + 2 FMNMX + 2 SHF  => 8 instr
            asm volatile("add.rn.f32x2 %0, %0, %8; add.rn.f32x2 %1, %1, %8; max.f32 %4, %4, %9; shf.l.wrap.b32 %6, %10, %6, 1;"
                         "add.rn.f32x2 %2, %2, %8; add.rn.f32x2 %3, %3, %8; max.f32 %5, %5, %9; shf.l.wrap.b32 %7, %10, %7, 1;"
                         : "+l"(p0), "+l"(p1), "+l"(p2), "+l"(p3), "+f"(x4), "+f"(x5), "+r"(u0), "+r"(u1) : "l"(cc), "f"(c0), "r"(__float_as_uint(c1)));
        } else if(OP == 13) {  // 6 FADD + 1 FMNMX + 1 SHF
            asm volatile("add.rn.f32 %0, %0, %8; add.rn.f32 %1, %1, %8; add.rn.f32 %2, %2, %8; max.f32 %6, %6, %8;"
                         "add.rn.f32 %3, %3, %8; add.rn.f32 %4, %4, %8; add.rn.f32 %5, %5, %8; shf.l.wrap.b32 %7, %9, %7, 1;"
                         : "+f"(x0), "+f"(x1), "+f"(x2), "+f"(x3), "+f"(x4), "+f"(x5), "+f"(x6), "+r"(u0) : "f"(c0), "r"(__float_as_uint(c1)));
        }
    }
    long long t1 = clock64();
    float2 q0 = *reinterpret_cast<float2*>(&p0), q1 = *reinterpret_cast<float2*>(&p1), q2 = *reinterpret_cast<float2*>(&p2), q3 = *reinterpret_cast<float2*>(&p3);
    out[blockIdx.x * blockDim.x + threadIdx.x] = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7 + q0.x + q1.y + q2.x + q3.y + __uint_as_float(u0 ^ u1 ^ u2 ^ u3);
    if(threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

template <int OP>
void run(const char* name, int instr_per_iter, double laneops_per_instr, int sms, float* out, long long* cyc) {
    bench<OP><<<sms, 512>>>(out, cyc, 1.0f, -0.5f);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    bench<OP><<<sms, 512>>>(out, cyc, 1.0f, -0.5f);
    cudaEventRecord(e1);
    cudaDeviceSynchronize();
    float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
    long long* h = new long long[sms];
    cudaMemcpy(h, cyc, sms * sizeof(long long), cudaMemcpyDeviceToHost);
    long long mx = 0; for(int i = 0; i < sms; ++i) mx = h[i] > mx ? h[i] : mx;
    double winstr = 16.0 * ITER * instr_per_iter;  // warp-instructions per SM
    printf("%-28s %8.3f warp-instr/clk/SM  %8.1f lane-ops/clk/SM  (cycles %lld, %.3f ms, %.0f MHz)\n", name,
           winstr / mx, winstr / mx * 32 * laneops_per_instr, mx, ms, mx / (ms * 1e3));
    delete[] h;
}

int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    int sms = p.multiProcessorCount;
    float* out; long long* cyc;
    cudaMalloc(&out, sms * 512 * sizeof(float)); cudaMalloc(&cyc, sms * sizeof(long long));
    printf("%s, %d SMs\n", p.name, sms);
    run<0>("FADD", 8, 1, sms, out, cyc);
    run<1>("FADD2", 8, 2, sms, out, cyc);
    run<2>("FMNMX", 8, 1, sms, out, cyc);
    run<3>("FMNMX3", 8, 1, sms, out, cyc);
    run<4>("SHF.L.W", 8, 1, sms, out, cyc);
    run<5>("FADD+FMNMX 1:1", 8, 1, sms, out, cyc);
    run<6>("FADD2+FMNMX 1:1", 8, 1.5, sms, out, cyc);
    run<7>("FADD+SHF 1:1", 8, 1, sms, out, cyc);
    run<8>("FSETP+SEL 1:1", 8, 1, sms, out, cyc);
    run<9>("SHFL.UP", 8, 1, sms, out, cyc);
    run<10>("2 LDS.128 + 6 FADD", 8, 1, sms, out, cyc);
    run<11>("LOP3", 8, 1, sms, out, cyc);
    run<12>("4 FADD2+2 FMNMX+2 SHF", 8, 1.5, sms, out, cyc);
    run<13>("6 FADD+1 FMNMX+1 SHF", 8, 1, sms, out, cyc);
    return 0;
}
