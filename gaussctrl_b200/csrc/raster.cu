// 3D-Gaussian rasterisation forward: projection, spherical harmonics, depth ordering, tile binning, compositing.
// Replaces gsplat 0.1.3's CUDA extension behind gc_model.py:140-154 (project_gaussians), :166 (spherical_harmonics),
// :174-186 and :191-202 (rasterize_gaussians).  fp32 SIMT, HBM/L2-bound.
//
// Tile binning (depth order, intersections, tile runs) lives in raster_bin.cu.
//
// Every fp32 expression of the projection kernel is an explicit tree of single IEEE operations (__fmul_rn /
// __fadd_rn, never contracted to FMA) mirroring oracle/gsplat_ref.py so the result is bit-exact.
#include "../../include/gaussctrl_b200.h"
#include "common.cuh"

namespace {

constexpr int BLOCK = 16;

__device__ __forceinline__ float mul(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ float add(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ float sub(float a, float b) { return __fsub_rn(a, b); }
__device__ __forceinline__ float rcp(float a) { return __fdiv_rn(1.0f, a); }
// ((a0*b0 + a1*b1) + a2*b2)
__device__ __forceinline__ float dot3(float a0, float b0, float a1, float b1, float a2, float b2) {
    return add(add(mul(a0, b0), mul(a1, b1)), mul(a2, b2));
}
__device__ __forceinline__ int f2i_sat(float v) { return __float2int_rz(v); }  // cvt.rzi.s32.f32 saturates, NaN -> 0
__device__ __forceinline__ int clampi(int v, int lo, int hi) { return min(hi, max(lo, v)); }

struct ProjConst {
    float vm[12];  // rows 0..2 of the view matrix
    float pm[16];
    float fx, fy, cx, cy, lim_x, lim_y, clip, glob_scale, half_w, half_h;
    int tbx, tby;
};

__device__ __forceinline__ void tile_bbox(float x, float y, float radius, int tbx, int tby, int& x0, int& x1, int& y0,
                                          int& y1) {
    const float blk = (float)BLOCK;
    const float tcx = __fdiv_rn(x, blk), tcy = __fdiv_rn(y, blk), tr = __fdiv_rn(radius, blk);
    x0 = clampi(f2i_sat(sub(tcx, tr)), 0, tbx);
    x1 = clampi(f2i_sat(add(add(tcx, tr), 1.0f)), 0, tbx);
    y0 = clampi(f2i_sat(sub(tcy, tr)), 0, tby);
    y1 = clampi(f2i_sat(add(add(tcy, tr), 1.0f)), 0, tby);
}

// projection of one Gaussian (gsplat project_gaussians_forward_kernel), shared by the seam kernel and the fused one
__device__ __forceinline__ void project_one(const ProjConst& P, int i, float px, float py, float pz, float s0, float s1,
                                            float s2, float w, float x, float y, float z, float* __restrict__ xys,
                                            float* __restrict__ depths, int32_t* __restrict__ radii,
                                            float* __restrict__ conics, int32_t* __restrict__ nth,
                                            float* __restrict__ cov3d) {
    const float* vm = P.vm;
    const float tx = add(dot3(vm[0], px, vm[1], py, vm[2], pz), vm[3]);
    const float ty = add(dot3(vm[4], px, vm[5], py, vm[6], pz), vm[7]);
    const float tz = add(dot3(vm[8], px, vm[9], py, vm[10], pz), vm[11]);
    bool valid = tz > P.clip;

    const float inv = rcp(__fsqrt_rn(add(add(add(mul(w, w), mul(x, x)), mul(y, y)), mul(z, z))));
    w = mul(w, inv);
    x = mul(x, inv);
    y = mul(y, inv);
    z = mul(z, inv);
    const float r00 = sub(1.0f, mul(2.0f, add(mul(y, y), mul(z, z))));
    const float r01 = mul(2.0f, sub(mul(x, y), mul(w, z)));
    const float r02 = mul(2.0f, add(mul(x, z), mul(w, y)));
    const float r10 = mul(2.0f, add(mul(x, y), mul(w, z)));
    const float r11 = sub(1.0f, mul(2.0f, add(mul(x, x), mul(z, z))));
    const float r12 = mul(2.0f, sub(mul(y, z), mul(w, x)));
    const float r20 = mul(2.0f, sub(mul(x, z), mul(w, y)));
    const float r21 = mul(2.0f, add(mul(y, z), mul(w, x)));
    const float r22 = sub(1.0f, mul(2.0f, add(mul(x, x), mul(y, y))));
    const float sx = mul(P.glob_scale, s0), sy = mul(P.glob_scale, s1), sz = mul(P.glob_scale, s2);
    const float m00 = mul(r00, sx), m01 = mul(r01, sy), m02 = mul(r02, sz);
    const float m10 = mul(r10, sx), m11 = mul(r11, sy), m12 = mul(r12, sz);
    const float m20 = mul(r20, sx), m21 = mul(r21, sy), m22 = mul(r22, sz);
    const float c00 = dot3(m00, m00, m01, m01, m02, m02);
    const float c01 = dot3(m00, m10, m01, m11, m02, m12);
    const float c02 = dot3(m00, m20, m01, m21, m02, m22);
    const float c11 = dot3(m10, m10, m11, m11, m12, m12);
    const float c12 = dot3(m10, m20, m11, m21, m12, m22);
    const float c22 = dot3(m20, m20, m21, m21, m22, m22);
    if (cov3d) {
        float* c = cov3d + 6ll * i;
        c[0] = c00;
        c[1] = c01;
        c[2] = c02;
        c[3] = c11;
        c[4] = c12;
        c[5] = c22;
    }
    const float tzs = valid ? tz : 1.0f;
    const float rz = rcp(tzs);
    const float txc = mul(tzs, fminf(fmaxf(mul(tx, rz), -P.lim_x), P.lim_x));
    const float tyc = mul(tzs, fminf(fmaxf(mul(ty, rz), -P.lim_y), P.lim_y));
    const float rz2 = mul(rz, rz);
    const float j00 = mul(P.fx, rz);
    const float j02 = mul(mul(-P.fx, txc), rz2);
    const float j11 = mul(P.fy, rz);
    const float j12 = mul(mul(-P.fy, tyc), rz2);
    const float t00 = add(mul(j00, vm[0]), mul(j02, vm[8]));
    const float t01 = add(mul(j00, vm[1]), mul(j02, vm[9]));
    const float t02 = add(mul(j00, vm[2]), mul(j02, vm[10]));
    const float t10 = add(mul(j11, vm[4]), mul(j12, vm[8]));
    const float t11 = add(mul(j11, vm[5]), mul(j12, vm[9]));
    const float t12 = add(mul(j11, vm[6]), mul(j12, vm[10]));
    const float v0x = dot3(t00, c00, t01, c01, t02, c02);
    const float v0y = dot3(t00, c01, t01, c11, t02, c12);
    const float v0z = dot3(t00, c02, t01, c12, t02, c22);
    const float v1x = dot3(t10, c00, t11, c01, t12, c02);
    const float v1y = dot3(t10, c01, t11, c11, t12, c12);
    const float v1z = dot3(t10, c02, t11, c12, t12, c22);
    const float a = add(dot3(v0x, t00, v0y, t01, v0z, t02), 0.3f);
    const float b = dot3(v0x, t10, v0y, t11, v0z, t12);
    const float c = add(dot3(v1x, t10, v1y, t11, v1z, t12), 0.3f);
    const float det = sub(mul(a, c), mul(b, b));
    valid = valid && (det != 0.0f);
    const float inv_det = rcp(det != 0.0f ? det : 1.0f);
    const float con0 = mul(c, inv_det), con1 = mul(-b, inv_det), con2 = mul(a, inv_det);
    const float b_mid = mul(0.5f, add(a, c));
    const float disc = __fsqrt_rn(fmaxf(sub(mul(b_mid, b_mid), det), 0.1f));
    const float v1 = add(b_mid, disc), v2 = sub(b_mid, disc);
    const float radius = ceilf(mul(3.0f, __fsqrt_rn(fmaxf(v1, v2))));
    const float* pm = P.pm;
    const float hx = add(dot3(pm[0], px, pm[1], py, pm[2], pz), pm[3]);
    const float hy = add(dot3(pm[4], px, pm[5], py, pm[6], pz), pm[7]);
    const float hw = add(dot3(pm[12], px, pm[13], py, pm[14], pz), pm[15]);
    const float rw = rcp(add(hw, 1e-6f));
    const float xs = sub(add(mul(P.half_w, mul(hx, rw)), P.cx), 0.5f);
    const float ys = sub(add(mul(P.half_h, mul(hy, rw)), P.cy), 0.5f);
    int x0, x1, y0, y1;
    tile_bbox(xs, ys, radius, P.tbx, P.tby, x0, x1, y0, y1);
    const int area = (x1 - x0) * (y1 - y0);
    valid = valid && area > 0;
    xys[2 * i] = valid ? xs : 0.f;
    xys[2 * i + 1] = valid ? ys : 0.f;
    depths[i] = valid ? tz : 0.f;
    radii[i] = valid ? f2i_sat(radius) : 0;
    conics[3 * i] = valid ? con0 : 0.f;
    conics[3 * i + 1] = valid ? con1 : 0.f;
    conics[3 * i + 2] = valid ? con2 : 0.f;
    nth[i] = valid ? area : 0;
}

__global__ void __launch_bounds__(256) project_kernel(const float* __restrict__ means, const float* __restrict__ scales,
                                                      const float* __restrict__ quats, const ProjConst P, int N,
                                                      float* __restrict__ xys, float* __restrict__ depths,
                                                      int32_t* __restrict__ radii, float* __restrict__ conics,
                                                      int32_t* __restrict__ nth, float* __restrict__ cov3d) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    project_one(P, i, means[3 * i], means[3 * i + 1], means[3 * i + 2], scales[3 * i], scales[3 * i + 1], scales[3 * i + 2],
                quats[4 * i], quats[4 * i + 1], quats[4 * i + 2], quats[4 * i + 3], xys, depths, radii, conics, nth, cov3d);
}

// ------------------------------------------------------------------------------------------ spherical harmonics
__constant__ float SH_C2[5] = {1.0925484305920792f, -1.0925484305920792f, 0.31539156525252005f, -1.0925484305920792f,
                               0.5462742152960396f};
__constant__ float SH_C3[7] = {-0.5900435899266435f, 2.890611442640554f, -0.4570457994644658f, 0.3731763325901154f,
                               -0.4570457994644658f, 1.445305721320277f,  -0.5900435899266435f};

__device__ __forceinline__ void sh_basis(int degree, float x, float y, float z, float* bas) {
    bas[0] = 0.28209479177387814f;
    if (degree < 1) return;
    const float C1 = 0.4886025119029199f;
    bas[1] = -C1 * y;
    bas[2] = C1 * z;
    bas[3] = -C1 * x;
    if (degree < 2) return;
    const float xx = x * x, xy = x * y, xz = x * z, yy = y * y, yz = y * z, zz = z * z;
    bas[4] = SH_C2[0] * xy;
    bas[5] = SH_C2[1] * yz;
    bas[6] = SH_C2[2] * (2.0f * zz - xx - yy);
    bas[7] = SH_C2[3] * xz;
    bas[8] = SH_C2[4] * (xx - yy);
    if (degree < 3) return;
    bas[9] = SH_C3[0] * y * (3.0f * xx - yy);
    bas[10] = SH_C3[1] * xy * z;
    bas[11] = SH_C3[2] * y * (4.0f * zz - xx - yy);
    bas[12] = SH_C3[3] * z * (2.0f * zz - 3.0f * xx - 3.0f * yy);
    bas[13] = SH_C3[4] * x * (4.0f * zz - xx - yy);
    bas[14] = SH_C3[5] * z * (xx - yy);
    bas[15] = SH_C3[6] * x * (xx - 3.0f * yy);
}

// one warp handles 32 Gaussians; coefficient rows (K*3 floats each) are read coalesced through shared memory
__global__ void __launch_bounds__(256) sh_fwd_kernel(int degree, int K, const float* __restrict__ dirs,
                                                     const float* __restrict__ coeffs, float* __restrict__ colors, int N) {
    extern __shared__ float s_co[];  // [8 warps][32][K*3 + 1]
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int row = K * 3, ldr = row + 1;
    float* sw = s_co + warp * 32 * ldr;
    const long long g0 = ((long long)blockIdx.x * 8 + warp) * 32;
    if (g0 >= N) return;
    const int cnt = (int)min(32ll, N - g0);
    const float* src = coeffs + g0 * row;
    for (int i = lane; i < cnt * row; i += 32) sw[(i / row) * ldr + i % row] = src[i];
    __syncwarp();
    if (lane < cnt) {
        const long long gi = g0 + lane;
        float bas[16];
        sh_basis(degree, dirs[3 * gi], dirs[3 * gi + 1], dirs[3 * gi + 2], bas);
        const int nb = (degree + 1) * (degree + 1);
        float r = 0.f, g = 0.f, b = 0.f;
        const float* c = sw + lane * ldr;
        for (int k = 0; k < nb; ++k) {
            r += bas[k] * c[3 * k];
            g += bas[k] * c[3 * k + 1];
            b += bas[k] * c[3 * k + 2];
        }
        colors[3 * gi] = r;
        colors[3 * gi + 1] = g;
        colors[3 * gi + 2] = b;
    }
}

// Fused eval-path front end of GaussCtrlModel.get_outputs (gc_model.py:138-167): ONE pass over the 236 B/Gaussian
// parameter record does exp(scales), quaternion normalisation, projection, view direction, SH colour (+0.5, clamp >= 0)
// and sigmoid(opacity).  A warp stages its 32 Gaussians' 45 SH-rest floats through shared memory (coalesced reads).
// Also writes (r, g, b, depth) as the 4-channel colour of the fused rgb+depth composite.
__global__ void __launch_bounds__(256) project_sh_fused_kernel(
    const float* __restrict__ means, const float* __restrict__ log_scales, const float* __restrict__ quats,
    const float* __restrict__ fdc, const float* __restrict__ frest, const float* __restrict__ opac_logit,
    const ProjConst P, float ox, float oy, float oz, int degree, int N, float* __restrict__ xys,
    float* __restrict__ depths, int32_t* __restrict__ radii, float* __restrict__ conics, int32_t* __restrict__ nth,
    float* __restrict__ rgbd, float* __restrict__ opac) {
    __shared__ float s_rest[8][32 * 45 + 1];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long g0 = ((long long)blockIdx.x * 8 + warp) * 32;
    if (g0 >= N) return;
    const int cnt = (int)min(32ll, N - g0);
    if (degree > 0) {
        const float* src = frest + g0 * 45;
        for (int i = lane; i < cnt * 45; i += 32) s_rest[warp][i] = src[i];
    }
    __syncwarp();
    if (lane >= cnt) return;
    const int i = (int)(g0 + lane);
    const float px = means[3 * i], py = means[3 * i + 1], pz = means[3 * i + 2];
    // quats / quats.norm() as the reference does before calling project_gaussians (gc_model.py:144)
    float qw = quats[4 * i], qx = quats[4 * i + 1], qy = quats[4 * i + 2], qz = quats[4 * i + 3];
    const float qn = __fsqrt_rn(add(add(add(mul(qw, qw), mul(qx, qx)), mul(qy, qy)), mul(qz, qz)));
    qw = __fdiv_rn(qw, qn);
    qx = __fdiv_rn(qx, qn);
    qy = __fdiv_rn(qy, qn);
    qz = __fdiv_rn(qz, qn);
    project_one(P, i, px, py, pz, expf(log_scales[3 * i]), expf(log_scales[3 * i + 1]), expf(log_scales[3 * i + 2]), qw,
                qx, qy, qz, xys, depths, radii, conics, nth, nullptr);
    // view direction from the camera origin, SH colour
    float dx = px - ox, dy = py - oy, dz = pz - oz;
    const float dn = sqrtf(dx * dx + dy * dy + dz * dz);
    dx /= dn;
    dy /= dn;
    dz /= dn;
    float bas[16];
    sh_basis(degree, dx, dy, dz, bas);
    float r = bas[0] * fdc[3 * i], g = bas[0] * fdc[3 * i + 1], b = bas[0] * fdc[3 * i + 2];
    const int nb = (degree + 1) * (degree + 1);
    const float* c = &s_rest[warp][lane * 45];
    for (int k = 1; k < nb; ++k) {
        r += bas[k] * c[3 * (k - 1)];
        g += bas[k] * c[3 * (k - 1) + 1];
        b += bas[k] * c[3 * (k - 1) + 2];
    }
    float4 o;
    o.x = fmaxf(r + 0.5f, 0.f);
    o.y = fmaxf(g + 0.5f, 0.f);
    o.z = fmaxf(b + 0.5f, 0.f);
    o.w = depths[i];
    reinterpret_cast<float4*>(rgbd)[i] = o;
    opac[i] = 1.f / (1.f + expf(-opac_logit[i]));
}

// ------------------------------------------------------------------------------------------ compositing
// FUSED: C == 4 channels are (r, g, b, depth) and the get_outputs epilogue (gc_model.py:187-204) is applied in place:
// out = rgb [H,W,3] clamped to <= 1, out_depth = depth / alpha (1000 where alpha == 0), out_alpha = 1 - T.
template <int C, bool FUSED>
__global__ void __launch_bounds__(256) rasterize_fwd_kernel(const float* __restrict__ xys,
                                                            const float* __restrict__ conics,
                                                            const float* __restrict__ colors,
                                                            const float* __restrict__ opac,
                                                            const int32_t* __restrict__ gids,
                                                            const int32_t* __restrict__ bins,
                                                            const int32_t* __restrict__ radii, int H, int W, int tbx,
                                                            float bg0, float bg1, float bg2, float bg3,
                                                            float* __restrict__ out, float* __restrict__ final_T,
                                                            int32_t* __restrict__ final_idx,
                                                            float* __restrict__ out_depth,
                                                            float* __restrict__ out_alpha,
                                                            const float* __restrict__ d_bg) {
    __shared__ float4 s_xyo[256];   // x, y, opacity, conic.x
    __shared__ float2 s_con[256];   // conic.y, conic.z
    __shared__ float s_thr[256];    // sigma above which alpha = o * exp(-sigma) is certainly < 1/255
    __shared__ float s_r2[256];     // squared distance from the centre beyond which that holds for every pixel
    __shared__ float s_col[256 * C];
    const int tile = blockIdx.y * tbx + blockIdx.x;
    // a warp owns an 8 x 4 pixel sub-block of the tile (compact, so whole Gaussians can be culled per warp)
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int sbx = blockIdx.x * BLOCK + (warp & 1) * 8, sby = blockIdx.y * BLOCK + (warp >> 1) * 4;
    const int px_i = sbx + (lane & 7), py_i = sby + (lane >> 3);
    const bool inside = px_i < W && py_i < H;
    const float px = (float)px_i, py = (float)py_i;
    const float rx0 = (float)sbx, rx1 = (float)(sbx + 7), ry0 = (float)sby, ry1 = (float)(sby + 3);
    const int start = bins[2 * tile], end = bins[2 * tile + 1];
    float T = 1.f;
    float acc[C];
#pragma unroll
    for (int c = 0; c < C; ++c) acc[c] = 0.f;
    int last = 0;
    bool done = !inside;
    for (int b0 = start; b0 < end; b0 += 256) {
        if (__syncthreads_and(done)) break;
        const int i = b0 + threadIdx.x;
        if (i < end) {
            const int g = gids[i];
            const float2 xy = reinterpret_cast<const float2*>(xys)[g];
            const float o = opac[g];
            s_xyo[threadIdx.x] = make_float4(xy.x, xy.y, o, conics[3 * g]);
            s_con[threadIdx.x] = make_float2(conics[3 * g + 1], conics[3 * g + 2]);
            // alpha < 1/255  <=>  sigma > ln(255 o); the margin (1e-3 relative in alpha) is four orders of magnitude above
            // the rounding of expf / logf, so every pair skipped here would have failed the exact test below as well:
            // results are bit-identical, the exponential is only evaluated near or inside the Gaussian's support
            const float thr = logf(255.f * o) + 1e-3f;
            s_thr[threadIdx.x] = thr;
            // sigma >= 0.5 * lambda_min(conic) * d^2 and lambda_min(conic) = 1 / lambda_max(cov2d) >= 9 / radius^2
            // (radius = ceil(3 sqrt(lambda_max)), project_one): beyond d^2 = thr * radius^2 / 4.5 no pixel passes the
            // exact test, so a warp whose sub-block is farther away skips the Gaussian without evaluating it
            float r2 = __int_as_float(0x7f800000);
            if (radii) {
                const float r = (float)radii[g];
                r2 = thr > 0.f ? thr * r * r * (1.0001f / 4.5f) + 1e-3f : -1.f;
            }
            s_r2[threadIdx.x] = r2;
#pragma unroll
            for (int c = 0; c < C; ++c) s_col[threadIdx.x * C + c] = colors[(long long)g * C + c];
        }
        __syncthreads();
        const int cnt = min(256, end - b0);
        for (int r0 = 0; r0 < cnt; r0 += 32) {
            if (__all_sync(0xffffffffu, done)) break;
            // which of these 32 Gaussians can touch this warp's sub-block at all (front-to-back order is kept below)
            const int jt = r0 + lane;
            bool hit = false;
            if (jt < cnt) {
                const float4 qt = s_xyo[jt];
                const float ddx = fmaxf(fmaxf(rx0 - qt.x, qt.x - rx1), 0.f), ddy = fmaxf(fmaxf(ry0 - qt.y, qt.y - ry1), 0.f);
                hit = !(ddx * ddx + ddy * ddy > s_r2[jt]);
            }
            unsigned m = __ballot_sync(0xffffffffu, hit);
            while (m) {
                const int j = r0 + __ffs(m) - 1;
                m &= m - 1;
                if (done) continue;
                const float4 q = s_xyo[j];
                const float2 cc = s_con[j];
                const float dx = q.x - px, dy = q.y - py;
                const float sigma = 0.5f * (q.w * dx * dx + cc.y * dy * dy) + cc.x * dx * dy;
                if (sigma > s_thr[j]) continue;
                const float alpha = fminf(0.999f, q.z * expf(-sigma));
                if (sigma < 0.f || alpha < (1.f / 255.f)) continue;
                const float nT = T * (1.f - alpha);
                if (nT <= 1e-4f) {
                    done = true;
                    continue;
                }
                const float vis = alpha * T;
#pragma unroll
                for (int c = 0; c < C; ++c) acc[c] += vis * s_col[j * C + c];
                T = nT;
                last = b0 + j;
            }
        }
    }
    if (inside) {
        const long long pid = (long long)py_i * W + px_i;
        float bg[4] = {bg0, bg1, bg2, bg3};
        if constexpr (FUSED) {
            bg[0] = d_bg[0];
            bg[1] = d_bg[1];
            bg[2] = d_bg[2];
            static_assert(C == 4, "fused rgb+depth epilogue needs (r, g, b, depth)");
            const float a = 1.f - T;
            out[pid * 3] = fminf(acc[0] + T * bg[0], 1.f);
            out[pid * 3 + 1] = fminf(acc[1] + T * bg[1], 1.f);
            out[pid * 3 + 2] = fminf(acc[2] + T * bg[2], 1.f);
            out_depth[pid] = a > 0.f ? (acc[3] + T * bg[3]) / a : 1000.f;
            out_alpha[pid] = a;
            if (final_T) final_T[pid] = T;
            if (final_idx) final_idx[pid] = last;
        } else {
#pragma unroll
            for (int c = 0; c < C; ++c) out[pid * C + c] = acc[c] + T * bg[c];
            final_T[pid] = T;
            final_idx[pid] = last;
        }
    }
}

__global__ void raster_finalize_kernel(const float* __restrict__ img4, const float* __restrict__ final_T,
                                       float* __restrict__ rgb, float* __restrict__ depth, float* __restrict__ alpha,
                                       int HW) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= HW) return;
    const float4 v = reinterpret_cast<const float4*>(img4)[i];
    const float a = 1.f - final_T[i];
    rgb[3 * i] = fminf(v.x, 1.f);
    rgb[3 * i + 1] = fminf(v.y, 1.f);
    rgb[3 * i + 2] = fminf(v.z, 1.f);
    depth[i] = a > 0.f ? v.w / a : 1000.f;
    alpha[i] = a;
}

}  // namespace

#define ST ((cudaStream_t)stream)

extern "C" int gcb_project_gaussians_fwd(const float* means3d, const float* scales, float glob_scale, const float* quats,
                                         const float* h_viewmat, const float* h_projmat, float fx, float fy, float cx,
                                         float cy, int img_h, int img_w, int tile_bx, int tile_by, float clip_thresh,
                                         int N, float* xys, float* depths, int32_t* radii, float* conics,
                                         int32_t* num_tiles_hit, float* cov3d, void* stream) {
    GCB_CHECK_ARG(means3d && scales && quats && h_viewmat && h_projmat, "null input");
    GCB_CHECK_ARG(xys && depths && radii && conics && num_tiles_hit, "null output");
    GCB_CHECK_ARG(N >= 0, "N < 0");
    if (N == 0) return GCB_OK;
    ProjConst P;
    for (int i = 0; i < 12; ++i) P.vm[i] = h_viewmat[i];
    for (int i = 0; i < 16; ++i) P.pm[i] = h_projmat[i];
    P.fx = fx;
    P.fy = fy;
    P.cx = cx;
    P.cy = cy;
    // float32(1.3) * float32(0.5 * W / fx), evaluated like the oracle
    P.lim_x = 1.3f * (float)(0.5 * (double)img_w / (double)fx);
    P.lim_y = 1.3f * (float)(0.5 * (double)img_h / (double)fy);
    P.clip = clip_thresh;
    P.glob_scale = glob_scale;
    P.half_w = 0.5f * (float)img_w;
    P.half_h = 0.5f * (float)img_h;
    P.tbx = tile_bx;
    P.tby = tile_by;
    project_kernel<<<gcb_cdiv(N, 256), 256, 0, ST>>>(means3d, scales, quats, P, N, xys, depths, radii, conics,
                                                     num_tiles_hit, cov3d);
    GCB_LAUNCH_CHECK();
    return GCB_OK;
}

static void fill_proj_const(ProjConst& P, const float* h_viewmat, const float* h_projmat, float fx, float fy, float cx,
                            float cy, int img_h, int img_w, int tile_bx, int tile_by, float clip_thresh, float glob_scale) {
    for (int i = 0; i < 12; ++i) P.vm[i] = h_viewmat[i];
    for (int i = 0; i < 16; ++i) P.pm[i] = h_projmat[i];
    P.fx = fx;
    P.fy = fy;
    P.cx = cx;
    P.cy = cy;
    P.lim_x = 1.3f * (float)(0.5 * (double)img_w / (double)fx);
    P.lim_y = 1.3f * (float)(0.5 * (double)img_h / (double)fy);
    P.clip = clip_thresh;
    P.glob_scale = glob_scale;
    P.half_w = 0.5f * (float)img_w;
    P.half_h = 0.5f * (float)img_h;
    P.tbx = tile_bx;
    P.tby = tile_by;
}

extern "C" int gcb_project_sh_fused_fwd(const float* means3d, const float* log_scales, const float* quats,
                                        const float* features_dc, const float* features_rest,
                                        const float* opacity_logits, const float* h_viewmat, const float* h_projmat,
                                        const float* h_cam_origin, float fx, float fy, float cx, float cy, int img_h,
                                        int img_w, int tile_bx, int tile_by, int sh_degree, int N, float* xys,
                                        float* depths, int32_t* radii, float* conics, int32_t* num_tiles_hit,
                                        float* rgbd, float* opac, void* stream) {
    GCB_CHECK_ARG(means3d && log_scales && quats && features_dc && opacity_logits && h_viewmat && h_projmat && h_cam_origin,
                  "null input");
    GCB_CHECK_ARG(sh_degree >= 0 && sh_degree <= 3 && (sh_degree == 0 || features_rest), "bad SH degree / features_rest");
    GCB_CHECK_ARG(xys && depths && radii && conics && num_tiles_hit && rgbd && opac, "null output");
    if (N == 0) return GCB_OK;
    ProjConst P;
    fill_proj_const(P, h_viewmat, h_projmat, fx, fy, cx, cy, img_h, img_w, tile_bx, tile_by, 0.01f, 1.0f);
    project_sh_fused_kernel<<<gcb_cdiv(N, 256), 256, 0, ST>>>(means3d, log_scales, quats, features_dc, features_rest,
                                                              opacity_logits, P, h_cam_origin[0], h_cam_origin[1],
                                                              h_cam_origin[2], sh_degree, N, xys, depths, radii, conics,
                                                              num_tiles_hit, rgbd, opac);
    GCB_LAUNCH_CHECK();
    return GCB_OK;
}

extern "C" int gcb_sh_fwd(int degree, int K, const float* viewdirs, const float* coeffs, float* colors, int N,
                          void* stream) {
    GCB_CHECK_ARG(viewdirs && coeffs && colors, "null pointer");
    GCB_CHECK_ARG(degree >= 0 && degree <= 3 && K >= (degree + 1) * (degree + 1) && K <= 16, "bad degree=%d / K=%d",
                  degree, K);
    if (N == 0) return GCB_OK;
    const size_t smem = (size_t)8 * 32 * (K * 3 + 1) * sizeof(float);
    static unsigned long long configured = 0;   // one bit per device ordinal
    if (gcb_first_use_on_device(configured)) {
        GCB_CUDA(cudaFuncSetAttribute(sh_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
    }
    sh_fwd_kernel<<<gcb_cdiv(N, 256), 256, smem, ST>>>(degree, K, viewdirs, coeffs, colors, N);
    GCB_LAUNCH_CHECK();
    return GCB_OK;
}

extern "C" int gcb_rasterize_fwd(const float* xys, const float* conics, const float* colors, const float* opacities,
                                 const int32_t* gaussian_ids, const int32_t* tile_bins, const int32_t* radii, int img_h,
                                 int img_w, int C, const float* h_background, float* out_img, float* final_T,
                                 int32_t* final_idx, void* stream) {
    GCB_CHECK_ARG(xys && conics && colors && opacities && gaussian_ids && tile_bins, "null input");
    GCB_CHECK_ARG(out_img && final_T && final_idx && h_background, "null output/background");
    const int tbx = gcb_cdiv(img_w, BLOCK), tby = gcb_cdiv(img_h, BLOCK);
    dim3 grid(tbx, tby);
    float bg[4] = {0.f, 0.f, 0.f, 0.f};
    for (int c = 0; c < C && c < 4; ++c) bg[c] = h_background[c];
    switch (C) {
        case 1:
            rasterize_fwd_kernel<1, false><<<grid, 256, 0, ST>>>(xys, conics, colors, opacities, gaussian_ids, tile_bins,
                                                                 radii, img_h, img_w, tbx, bg[0], bg[1], bg[2], bg[3], out_img,
                                                                 final_T, final_idx, nullptr, nullptr, nullptr);
            break;
        case 3:
            rasterize_fwd_kernel<3, false><<<grid, 256, 0, ST>>>(xys, conics, colors, opacities, gaussian_ids, tile_bins,
                                                                 radii, img_h, img_w, tbx, bg[0], bg[1], bg[2], bg[3], out_img,
                                                                 final_T, final_idx, nullptr, nullptr, nullptr);
            break;
        case 4:
            rasterize_fwd_kernel<4, false><<<grid, 256, 0, ST>>>(xys, conics, colors, opacities, gaussian_ids, tile_bins,
                                                                 radii, img_h, img_w, tbx, bg[0], bg[1], bg[2], bg[3], out_img,
                                                                 final_T, final_idx, nullptr, nullptr, nullptr);
            break;
        default:
            gcb_set_error("rasterize: C=%d not built (1, 3, 4)", C);
            return GCB_ERR_UNSUPPORTED;
    }
    GCB_LAUNCH_CHECK();
    return GCB_OK;
}

extern "C" int gcb_rasterize_rgbd_fwd(const float* xys, const float* conics, const float* rgbd, const float* opacities,
                                      const int32_t* gaussian_ids, const int32_t* tile_bins, const int32_t* radii,
                                      int img_h, int img_w, const float* d_background3, float* out_rgb, float* out_depth,
                                      float* out_alpha, void* stream) {
    GCB_CHECK_ARG(xys && conics && rgbd && opacities && gaussian_ids && tile_bins && d_background3, "null input");
    GCB_CHECK_ARG(out_rgb && out_depth && out_alpha, "null output");
    const int tbx = gcb_cdiv(img_w, BLOCK), tby = gcb_cdiv(img_h, BLOCK);
    dim3 grid(tbx, tby);
    rasterize_fwd_kernel<4, true><<<grid, 256, 0, ST>>>(xys, conics, rgbd, opacities, gaussian_ids, tile_bins, radii, img_h,
                                                        img_w, tbx, 0.f, 0.f, 0.f, 0.f, out_rgb, nullptr, nullptr, out_depth,
                                                        out_alpha, d_background3);
    GCB_LAUNCH_CHECK();
    return GCB_OK;
}

extern "C" int gcb_raster_finalize(const float* img4, const float* final_T, float* rgb, float* depth, float* alpha,
                                   int HW, void* stream) {
    GCB_CHECK_ARG(img4 && final_T && rgb && depth && alpha, "null pointer");
    raster_finalize_kernel<<<gcb_cdiv(HW, 256), 256, 0, ST>>>(img4, final_T, rgb, depth, alpha, HW);
    GCB_LAUNCH_CHECK();
    return GCB_OK;
}

// ------------------------------------------------------------------------------------------ batched eval renders
// V eval-mode views of one scene in ONE call (render_reverse renders every training view, gc_pipeline.py:126-133): the
// host loop over fused project+SH -> binning -> fused rgb+depth composite runs here instead of in the interpreter
// (~13 launches per view; from Python the loop was launch-overhead-bound at 0.36 ms per view, above the GPU time).
// The views are stream-ordered, so they share one set of intermediates; every view has its own outputs and its own
// (M, overflow) pair.  Nothing synchronises.
namespace {
inline size_t a256(size_t v) { return (v + 255) & ~(size_t)255; }
}  // namespace

extern "C" size_t gcb_bin_gaussians_workspace_bytes(int N, long long isect_capacity, int tile_bx, int tile_by);
extern "C" int gcb_bin_gaussians(const float* xys, const float* depths, const int32_t* radii, const int32_t* num_tiles_hit,
                                 int N, int tile_bx, int tile_by, long long isect_capacity, int32_t* gaussian_ids,
                                 int32_t* tile_bins, int32_t* isect_count, int64_t* isect_keys, void* workspace,
                                 size_t workspace_bytes, void* stream);

extern "C" size_t gcb_render_eval_batch_workspace_bytes(int N, long long isect_capacity, int img_h, int img_w) {
    if (N <= 0 || isect_capacity <= 0 || img_h <= 0 || img_w <= 0) return 0;
    const int tbx = gcb_cdiv(img_w, BLOCK), tby = gcb_cdiv(img_h, BLOCK);
    return a256((size_t)N * 8) + 3 * a256((size_t)N * 4) + a256((size_t)N * 12) + a256((size_t)N * 16) + a256((size_t)N * 4) +
           a256((size_t)isect_capacity * 4) + a256((size_t)tbx * tby * 8) +
           gcb_bin_gaussians_workspace_bytes(N, isect_capacity, tbx, tby);
}

extern "C" int gcb_render_eval_batch(const float* means3d, const float* log_scales, const float* quats,
                                     const float* features_dc, const float* features_rest, const float* opacity_logits,
                                     int N, int sh_degree, int V, const float* h_viewmats, const float* h_projmats,
                                     const float* h_cam_origins, const float* h_intrinsics, int img_h, int img_w,
                                     const float* d_background3, long long isect_capacity, float* out_rgb, float* out_depth,
                                     float* out_alpha, int32_t* isect_counts, void* workspace, size_t workspace_bytes,
                                     void* stream) {
    GCB_CHECK_ARG(h_viewmats && h_projmats && h_cam_origins && h_intrinsics && d_background3, "null camera / background");
    GCB_CHECK_ARG(out_rgb && out_depth && out_alpha && isect_counts && workspace, "null output / workspace");
    GCB_CHECK_ARG(V >= 0 && N > 0, "bad V=%d / N=%d", V, N);
    const size_t need = gcb_render_eval_batch_workspace_bytes(N, isect_capacity, img_h, img_w);
    if (workspace_bytes < need || need == 0) {
        gcb_set_error("render batch workspace too small: %zu < %zu", workspace_bytes, need);
        return GCB_ERR_WORKSPACE;
    }
    const int tbx = gcb_cdiv(img_w, BLOCK), tby = gcb_cdiv(img_h, BLOCK);
    char* w = (char*)workspace;
    auto take = [&](size_t bytes) {
        char* p = w;
        w += a256(bytes);
        return p;
    };
    float* xys = (float*)take((size_t)N * 8);
    float* depths = (float*)take((size_t)N * 4);
    int32_t* radii = (int32_t*)take((size_t)N * 4);
    int32_t* nth = (int32_t*)take((size_t)N * 4);
    float* conics = (float*)take((size_t)N * 12);
    float* rgbd = (float*)take((size_t)N * 16);
    float* opac = (float*)take((size_t)N * 4);
    int32_t* gids = (int32_t*)take((size_t)isect_capacity * 4);
    int32_t* bins = (int32_t*)take((size_t)tbx * tby * 8);
    const size_t bin_bytes = gcb_bin_gaussians_workspace_bytes(N, isect_capacity, tbx, tby);
    const size_t px = (size_t)img_h * img_w;
    for (int v = 0; v < V; ++v) {
        const float* in = h_intrinsics + 4 * v;
        int rc = gcb_project_sh_fused_fwd(means3d, log_scales, quats, features_dc, features_rest, opacity_logits,
                                          h_viewmats + 16 * v, h_projmats + 16 * v, h_cam_origins + 3 * v, in[0], in[1], in[2],
                                          in[3], img_h, img_w, tbx, tby, sh_degree, N, xys, depths, radii, conics, nth, rgbd,
                                          opac, stream);
        if (rc != GCB_OK) return rc;
        rc = gcb_bin_gaussians(xys, depths, radii, nth, N, tbx, tby, isect_capacity, gids, bins, isect_counts + 2 * v, nullptr,
                               w, bin_bytes, stream);
        if (rc != GCB_OK) return rc;
        rc = gcb_rasterize_rgbd_fwd(xys, conics, rgbd, opac, gids, bins, radii, img_h, img_w, d_background3,
                                    out_rgb + 3 * px * v, out_depth + px * v, out_alpha + px * v, stream);
        if (rc != GCB_OK) return rc;
    }
    return GCB_OK;
}
