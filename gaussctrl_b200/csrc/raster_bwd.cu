// Backward of the rasteriser (the 3DGS fine-tune step that follows the edit: gc_trainer.py:257-301 -> loss.backward()
// through gsplat's rasterize_gaussians / project_gaussians / spherical_harmonics).
// Gradients are the exact derivatives of the forward in raster.cu (checked against autograd of oracle/gsplat_ref.py):
//   rasterize_bwd : back-to-front per-pixel traversal of each tile, warp-shuffle reduction, one atomicAdd per warp
//   project_bwd   : (v_xy, v_depth, v_conic) -> (v_means3d, v_scales, v_quats), one thread per Gaussian
//   sh_bwd        : v_coeffs = basis(viewdir) (x) v_color
// v_conic here is the true gradient w.r.t. the stored conic (a, b, c) with sigma = 0.5(a dx^2 + c dy^2) + b dx dy.
#include "../../include/gaussctrl_b200.h"
#include "common.cuh"

namespace {

constexpr int BLOCK = 16;

template <int C>
__global__ void __launch_bounds__(256) rasterize_bwd_kernel(
    const float* __restrict__ xys, const float* __restrict__ conics, const float* __restrict__ colors,
    const float* __restrict__ opac, const int32_t* __restrict__ gids, const int32_t* __restrict__ bins, int H, int W,
    int tbx, float bg0, float bg1, float bg2, float bg3, const float* __restrict__ final_T,
    const int32_t* __restrict__ final_idx, const float* __restrict__ v_out, const float* __restrict__ v_out_alpha,
    float* __restrict__ v_xy, float* __restrict__ v_conic, float* __restrict__ v_colors, float* __restrict__ v_opacity) {
    __shared__ int s_id[256];
    __shared__ float4 s_xyo[256];  // x, y, opacity, conic.x
    __shared__ float2 s_con[256];  // conic.y, conic.z
    __shared__ float s_col[256 * C];
    const int tile = blockIdx.y * tbx + blockIdx.x;
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4, lane = threadIdx.x & 31;
    const int px_i = blockIdx.x * BLOCK + tx, py_i = blockIdx.y * BLOCK + ty;
    const bool inside = px_i < W && py_i < H;
    const float px = (float)px_i, py = (float)py_i;
    const long long pid = (long long)py_i * W + px_i;
    const int start = bins[2 * tile], end = bins[2 * tile + 1];
    if (end <= start) return;
    const float T_final = inside ? final_T[pid] : 0.f;
    const int bin_final = inside ? final_idx[pid] : -1;
    float T = T_final;
    float vo[C], buffer[C];
    const float bg[4] = {bg0, bg1, bg2, bg3};
    float bg_dot = 0.f;
#pragma unroll
    for (int c = 0; c < C; ++c) {
        vo[c] = inside ? v_out[pid * C + c] : 0.f;
        buffer[c] = 0.f;
        bg_dot += bg[c] * vo[c];
    }
    const float voa = (inside && v_out_alpha) ? v_out_alpha[pid] : 0.f;
    const int nbatch = (end - start + 255) / 256;
    for (int bi = nbatch - 1; bi >= 0; --bi) {
        const int b0 = start + bi * 256;
        __syncthreads();
        const int i = b0 + threadIdx.x;
        if (i < end) {
            const int g = gids[i];
            s_id[threadIdx.x] = g;
            const float2 xy = reinterpret_cast<const float2*>(xys)[g];
            s_xyo[threadIdx.x] = make_float4(xy.x, xy.y, opac[g], conics[3 * g]);
            s_con[threadIdx.x] = make_float2(conics[3 * g + 1], conics[3 * g + 2]);
#pragma unroll
            for (int c = 0; c < C; ++c) s_col[threadIdx.x * C + c] = colors[(long long)g * C + c];
        }
        __syncthreads();
        const int cnt = min(256, end - b0);
        for (int j = cnt - 1; j >= 0; --j) {
            const int idx = b0 + j;
            bool valid = inside && idx <= bin_final;
            float dx = 0.f, dy = 0.f, vis = 0.f, alpha = 0.f, o = 0.f;
            float4 q = make_float4(0.f, 0.f, 0.f, 0.f);
            float2 cc = make_float2(0.f, 0.f);
            if (valid) {
                q = s_xyo[j];
                cc = s_con[j];
                o = q.z;
                dx = q.x - px;
                dy = q.y - py;
                const float sigma = 0.5f * (q.w * dx * dx + cc.y * dy * dy) + cc.x * dx * dy;
                vis = expf(-sigma);
                alpha = fminf(0.999f, o * vis);
                if (sigma < 0.f || alpha < (1.f / 255.f)) valid = false;
            }
            if (!__any_sync(0xffffffffu, valid)) continue;
            float g_rgb[C];
            float g_xy0 = 0.f, g_xy1 = 0.f, g_c0 = 0.f, g_c1 = 0.f, g_c2 = 0.f, g_o = 0.f;
#pragma unroll
            for (int c = 0; c < C; ++c) g_rgb[c] = 0.f;
            if (valid) {
                const float ra = 1.f / (1.f - alpha);
                T *= ra;  // transmittance in front of this Gaussian
                const float fac = alpha * T;
                float v_alpha = 0.f;
#pragma unroll
                for (int c = 0; c < C; ++c) {
                    const float col = s_col[j * C + c];
                    g_rgb[c] = fac * vo[c];
                    v_alpha += (col * T - buffer[c] * ra) * vo[c];
                    buffer[c] += col * fac;
                }
                v_alpha += T_final * ra * voa;
                v_alpha -= T_final * ra * bg_dot;
                if (o * vis <= 0.999f) {
                    const float v_sigma = -o * vis * v_alpha;
                    g_c0 = 0.5f * v_sigma * dx * dx;
                    g_c1 = v_sigma * dx * dy;
                    g_c2 = 0.5f * v_sigma * dy * dy;
                    g_xy0 = v_sigma * (q.w * dx + cc.x * dy);
                    g_xy1 = v_sigma * (cc.x * dx + cc.y * dy);
                    g_o = vis * v_alpha;
                }
            }
#pragma unroll
            for (int c = 0; c < C; ++c) g_rgb[c] = warp_sum(g_rgb[c]);
            g_xy0 = warp_sum(g_xy0);
            g_xy1 = warp_sum(g_xy1);
            g_c0 = warp_sum(g_c0);
            g_c1 = warp_sum(g_c1);
            g_c2 = warp_sum(g_c2);
            g_o = warp_sum(g_o);
            if (lane == 0) {
                const int g = s_id[j];
#pragma unroll
                for (int c = 0; c < C; ++c) atomicAdd(&v_colors[(long long)g * C + c], g_rgb[c]);
                atomicAdd(&v_xy[2 * g], g_xy0);
                atomicAdd(&v_xy[2 * g + 1], g_xy1);
                atomicAdd(&v_conic[3 * g], g_c0);
                atomicAdd(&v_conic[3 * g + 1], g_c1);
                atomicAdd(&v_conic[3 * g + 2], g_c2);
                atomicAdd(&v_opacity[g], g_o);
            }
        }
    }
}

struct ProjBwdConst {
    float vm[12];
    float pm[16];
    float fx, fy, lim_x, lim_y, glob_scale, half_w, half_h;
};

__global__ void __launch_bounds__(256) project_bwd_kernel(
    const float* __restrict__ means, const float* __restrict__ scales, const float* __restrict__ quats,
    const ProjBwdConst P, const int32_t* __restrict__ radii, const float* __restrict__ v_xy,
    const float* __restrict__ v_depth, const float* __restrict__ v_conic, int N, float* __restrict__ v_means,
    float* __restrict__ v_scales, float* __restrict__ v_quats) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    float vp[3] = {0.f, 0.f, 0.f}, vs[3] = {0.f, 0.f, 0.f}, vq[4] = {0.f, 0.f, 0.f, 0.f};
    if (radii[i] > 0) {
        const float* vm = P.vm;
        const float* pm = P.pm;
        const float p[3] = {means[3 * i], means[3 * i + 1], means[3 * i + 2]};
        const float Wm[3][3] = {{vm[0], vm[1], vm[2]}, {vm[4], vm[5], vm[6]}, {vm[8], vm[9], vm[10]}};
        const float tx = Wm[0][0] * p[0] + Wm[0][1] * p[1] + Wm[0][2] * p[2] + vm[3];
        const float ty = Wm[1][0] * p[0] + Wm[1][1] * p[1] + Wm[1][2] * p[2] + vm[7];
        const float tz = Wm[2][0] * p[0] + Wm[2][1] * p[1] + Wm[2][2] * p[2] + vm[11];
        // rotation from the normalised quaternion, M = R S, V = M M^T
        float qw = quats[4 * i], qx = quats[4 * i + 1], qy = quats[4 * i + 2], qz = quats[4 * i + 3];
        const float qn = sqrtf(qw * qw + qx * qx + qy * qy + qz * qz), iqn = 1.f / qn;
        const float w = qw * iqn, x = qx * iqn, y = qy * iqn, z = qz * iqn;
        const float R[3][3] = {{1.f - 2.f * (y * y + z * z), 2.f * (x * y - w * z), 2.f * (x * z + w * y)},
                               {2.f * (x * y + w * z), 1.f - 2.f * (x * x + z * z), 2.f * (y * z - w * x)},
                               {2.f * (x * z - w * y), 2.f * (y * z + w * x), 1.f - 2.f * (x * x + y * y)}};
        const float s[3] = {P.glob_scale * scales[3 * i], P.glob_scale * scales[3 * i + 1], P.glob_scale * scales[3 * i + 2]};
        float M[3][3], V[3][3];
#pragma unroll
        for (int r = 0; r < 3; ++r)
#pragma unroll
            for (int c = 0; c < 3; ++c) M[r][c] = R[r][c] * s[c];
#pragma unroll
        for (int r = 0; r < 3; ++r)
#pragma unroll
            for (int c = 0; c < 3; ++c) V[r][c] = M[r][0] * M[c][0] + M[r][1] * M[c][1] + M[r][2] * M[c][2];
        // J, T = J W
        const float rz = 1.f / tz, rz2 = rz * rz;
        const float ux = tx * rz, uy = ty * rz;
        const bool cxm = fabsf(ux) <= P.lim_x, cym = fabsf(uy) <= P.lim_y;
        const float uxc = fminf(fmaxf(ux, -P.lim_x), P.lim_x), uyc = fminf(fmaxf(uy, -P.lim_y), P.lim_y);
        const float txc = tz * uxc, tyc = tz * uyc;
        const float j00 = P.fx * rz, j02 = -P.fx * txc * rz2, j11 = P.fy * rz, j12 = -P.fy * tyc * rz2;
        float T[2][3];
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            T[0][c] = j00 * Wm[0][c] + j02 * Wm[2][c];
            T[1][c] = j11 * Wm[1][c] + j12 * Wm[2][c];
        }
        float TV[2][3];
#pragma unroll
        for (int r = 0; r < 2; ++r)
#pragma unroll
            for (int c = 0; c < 3; ++c) TV[r][c] = T[r][0] * V[0][c] + T[r][1] * V[1][c] + T[r][2] * V[2][c];
        const float a = TV[0][0] * T[0][0] + TV[0][1] * T[0][1] + TV[0][2] * T[0][2] + 0.3f;
        const float b = TV[0][0] * T[1][0] + TV[0][1] * T[1][1] + TV[0][2] * T[1][2];
        const float c_ = TV[1][0] * T[1][0] + TV[1][1] * T[1][1] + TV[1][2] * T[1][2] + 0.3f;
        const float det = a * c_ - b * b, id = 1.f / det, id2 = id * id;
        const float g0 = v_conic[3 * i], g1 = v_conic[3 * i + 1], g2 = v_conic[3 * i + 2];
        const float v_a = g0 * (-c_ * c_ * id2) + g1 * (b * c_ * id2) + g2 * (id - a * c_ * id2);
        const float v_b = g0 * (2.f * b * c_ * id2) + g1 * (-id - 2.f * b * b * id2) + g2 * (2.f * a * b * id2);
        const float v_c = g0 * (id - a * c_ * id2) + g1 * (a * b * id2) + g2 * (-a * a * id2);
        const float G[2][2] = {{v_a, 0.5f * v_b}, {0.5f * v_b, v_c}};
        // v_Vfull = T^T G T ; v_T = 2 G (T V)
        float GT[2][3], vV[3][3], vT[2][3];
#pragma unroll
        for (int r = 0; r < 2; ++r)
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                GT[r][c] = G[r][0] * T[0][c] + G[r][1] * T[1][c];
                vT[r][c] = 2.f * (G[r][0] * TV[0][c] + G[r][1] * TV[1][c]);
            }
#pragma unroll
        for (int r = 0; r < 3; ++r)
#pragma unroll
            for (int c = 0; c < 3; ++c) vV[r][c] = T[0][r] * GT[0][c] + T[1][r] * GT[1][c];
        // v_J = v_T W^T (only the four non-zero entries of J matter)
        const float v_j00 = vT[0][0] * Wm[0][0] + vT[0][1] * Wm[0][1] + vT[0][2] * Wm[0][2];
        const float v_j02 = vT[0][0] * Wm[2][0] + vT[0][1] * Wm[2][1] + vT[0][2] * Wm[2][2];
        const float v_j11 = vT[1][0] * Wm[1][0] + vT[1][1] * Wm[1][1] + vT[1][2] * Wm[1][2];
        const float v_j12 = vT[1][0] * Wm[2][0] + vT[1][1] * Wm[2][1] + vT[1][2] * Wm[2][2];
        const float v_rz = P.fx * v_j00 + P.fy * v_j11 - 2.f * P.fx * txc * rz * v_j02 - 2.f * P.fy * tyc * rz * v_j12;
        const float v_txc = -P.fx * rz2 * v_j02, v_tyc = -P.fy * rz2 * v_j12;
        float vt[3];
        vt[0] = cxm ? v_txc : 0.f;
        vt[1] = cym ? v_tyc : 0.f;
        vt[2] = -rz2 * v_rz + v_depth[i] + (cxm ? 0.f : uxc * v_txc) + (cym ? 0.f : uyc * v_tyc);
#pragma unroll
        for (int c = 0; c < 3; ++c) vp[c] = vt[0] * Wm[0][c] + vt[1] * Wm[1][c] + vt[2] * Wm[2][c];
        // pixel position
        const float hx = pm[0] * p[0] + pm[1] * p[1] + pm[2] * p[2] + pm[3];
        const float hy = pm[4] * p[0] + pm[5] * p[1] + pm[6] * p[2] + pm[7];
        const float hw = pm[12] * p[0] + pm[13] * p[1] + pm[14] * p[2] + pm[15];
        const float rw = 1.f / (hw + 1e-6f);
        const float gx = v_xy[2 * i], gy = v_xy[2 * i + 1];
        const float v_hx = gx * P.half_w * rw, v_hy = gy * P.half_h * rw;
        const float v_hw = -(gx * P.half_w * hx + gy * P.half_h * hy) * rw * rw;
#pragma unroll
        for (int c = 0; c < 3; ++c) vp[c] += v_hx * pm[c] + v_hy * pm[4 + c] + v_hw * pm[12 + c];
        // v_M = 2 v_Vfull M ; v_s, v_R
        float vM[3][3], vR[3][3];
#pragma unroll
        for (int r = 0; r < 3; ++r)
#pragma unroll
            for (int c = 0; c < 3; ++c) vM[r][c] = 2.f * (vV[r][0] * M[0][c] + vV[r][1] * M[1][c] + vV[r][2] * M[2][c]);
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            vs[c] = P.glob_scale * (R[0][c] * vM[0][c] + R[1][c] * vM[1][c] + R[2][c] * vM[2][c]);
#pragma unroll
            for (int r = 0; r < 3; ++r) vR[r][c] = vM[r][c] * s[c];
        }
        const float v_w = 2.f * (x * (vR[2][1] - vR[1][2]) + y * (vR[0][2] - vR[2][0]) + z * (vR[1][0] - vR[0][1]));
        const float v_x = 2.f * (-2.f * x * (vR[1][1] + vR[2][2]) + y * (vR[1][0] + vR[0][1]) + z * (vR[2][0] + vR[0][2]) +
                                 w * (vR[2][1] - vR[1][2]));
        const float v_y = 2.f * (x * (vR[1][0] + vR[0][1]) - 2.f * y * (vR[0][0] + vR[2][2]) + z * (vR[2][1] + vR[1][2]) +
                                 w * (vR[0][2] - vR[2][0]));
        const float v_z = 2.f * (x * (vR[2][0] + vR[0][2]) + y * (vR[2][1] + vR[1][2]) - 2.f * z * (vR[0][0] + vR[1][1]) +
                                 w * (vR[1][0] - vR[0][1]));
        const float dotq = w * v_w + x * v_x + y * v_y + z * v_z;
        vq[0] = (v_w - w * dotq) * iqn;
        vq[1] = (v_x - x * dotq) * iqn;
        vq[2] = (v_y - y * dotq) * iqn;
        vq[3] = (v_z - z * dotq) * iqn;
    }
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        v_means[3 * i + c] = vp[c];
        v_scales[3 * i + c] = vs[c];
    }
#pragma unroll
    for (int c = 0; c < 4; ++c) v_quats[4 * i + c] = vq[c];
}

__constant__ float BSH_C2[5] = {1.0925484305920792f, -1.0925484305920792f, 0.31539156525252005f, -1.0925484305920792f,
                                0.5462742152960396f};
__constant__ float BSH_C3[7] = {-0.5900435899266435f, 2.890611442640554f, -0.4570457994644658f, 0.3731763325901154f,
                                -0.4570457994644658f, 1.445305721320277f,  -0.5900435899266435f};

__global__ void sh_bwd_kernel(int degree, int K, const float* __restrict__ dirs, const float* __restrict__ v_colors,
                              float* __restrict__ v_coeffs, int N) {
    const long long total = (long long)N * K;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const long long g = i / K;
        const int k = (int)(i % K);
        const float x = dirs[3 * g], y = dirs[3 * g + 1], z = dirs[3 * g + 2];
        const float xx = x * x, xy = x * y, xz = x * z, yy = y * y, yz = y * z, zz = z * z;
        float bas = 0.f;
        const int nb = (degree + 1) * (degree + 1);
        if (k < nb) {
            switch (k) {
                case 0: bas = 0.28209479177387814f; break;
                case 1: bas = -0.4886025119029199f * y; break;
                case 2: bas = 0.4886025119029199f * z; break;
                case 3: bas = -0.4886025119029199f * x; break;
                case 4: bas = BSH_C2[0] * xy; break;
                case 5: bas = BSH_C2[1] * yz; break;
                case 6: bas = BSH_C2[2] * (2.f * zz - xx - yy); break;
                case 7: bas = BSH_C2[3] * xz; break;
                case 8: bas = BSH_C2[4] * (xx - yy); break;
                case 9: bas = BSH_C3[0] * y * (3.f * xx - yy); break;
                case 10: bas = BSH_C3[1] * xy * z; break;
                case 11: bas = BSH_C3[2] * y * (4.f * zz - xx - yy); break;
                case 12: bas = BSH_C3[3] * z * (2.f * zz - 3.f * xx - 3.f * yy); break;
                case 13: bas = BSH_C3[4] * x * (4.f * zz - xx - yy); break;
                case 14: bas = BSH_C3[5] * z * (xx - yy); break;
                default: bas = BSH_C3[6] * x * (xx - 3.f * yy); break;
            }
        }
        v_coeffs[i * 3] = bas * v_colors[3 * g];
        v_coeffs[i * 3 + 1] = bas * v_colors[3 * g + 1];
        v_coeffs[i * 3 + 2] = bas * v_colors[3 * g + 2];
    }
}

}  // namespace

#define ST ((cudaStream_t)stream)

extern "C" int gcb_rasterize_bwd(const float* xys, const float* conics, const float* colors, const float* opacities,
                                 const int32_t* gaussian_ids, const int32_t* tile_bins, int img_h, int img_w, int C,
                                 const float* h_background, const float* final_T, const int32_t* final_idx,
                                 const float* v_out, const float* v_out_alpha, float* v_xy, float* v_conic,
                                 float* v_colors, float* v_opacity, void* stream) {
    GCB_CHECK_ARG(xys && conics && colors && opacities && gaussian_ids && tile_bins && final_T && final_idx && v_out,
                  "null input");
    GCB_CHECK_ARG(v_xy && v_conic && v_colors && v_opacity && h_background, "null output/background");
    const int tbx = gcb_cdiv(img_w, BLOCK), tby = gcb_cdiv(img_h, BLOCK);
    dim3 grid(tbx, tby);
    float bg[4] = {0.f, 0.f, 0.f, 0.f};
    for (int c = 0; c < C && c < 4; ++c) bg[c] = h_background[c];
#define GCB_RB(CC)                                                                                                      \
    rasterize_bwd_kernel<CC><<<grid, 256, 0, ST>>>(xys, conics, colors, opacities, gaussian_ids, tile_bins, img_h, img_w, \
                                                  tbx, bg[0], bg[1], bg[2], bg[3], final_T, final_idx, v_out,          \
                                                  v_out_alpha, v_xy, v_conic, v_colors, v_opacity)
    switch (C) {
        case 1: GCB_RB(1); break;
        case 3: GCB_RB(3); break;
        case 4: GCB_RB(4); break;
        default:
            gcb_set_error("rasterize backward: C=%d not built (1, 3, 4)", C);
            return GCB_ERR_UNSUPPORTED;
    }
#undef GCB_RB
    GCB_LAUNCH_CHECK();
    return GCB_OK;
}

extern "C" int gcb_project_gaussians_bwd(const float* means3d, const float* scales, float glob_scale, const float* quats,
                                         const float* h_viewmat, const float* h_projmat, float fx, float fy, float cx,
                                         float cy, int img_h, int img_w, const int32_t* radii, const float* v_xy,
                                         const float* v_depth, const float* v_conic, int N, float* v_means3d,
                                         float* v_scales, float* v_quats, void* stream) {
    (void)cx;
    (void)cy;
    GCB_CHECK_ARG(means3d && scales && quats && h_viewmat && h_projmat && radii && v_xy && v_depth && v_conic, "null input");
    GCB_CHECK_ARG(v_means3d && v_scales && v_quats, "null output");
    if (N == 0) return GCB_OK;
    ProjBwdConst P;
    for (int i = 0; i < 12; ++i) P.vm[i] = h_viewmat[i];
    for (int i = 0; i < 16; ++i) P.pm[i] = h_projmat[i];
    P.fx = fx;
    P.fy = fy;
    P.lim_x = 1.3f * (float)(0.5 * (double)img_w / (double)fx);
    P.lim_y = 1.3f * (float)(0.5 * (double)img_h / (double)fy);
    P.glob_scale = glob_scale;
    P.half_w = 0.5f * (float)img_w;
    P.half_h = 0.5f * (float)img_h;
    project_bwd_kernel<<<gcb_cdiv(N, 256), 256, 0, ST>>>(means3d, scales, quats, P, radii, v_xy, v_depth, v_conic, N,
                                                         v_means3d, v_scales, v_quats);
    GCB_LAUNCH_CHECK();
    return GCB_OK;
}

extern "C" int gcb_sh_bwd(int degree, int K, const float* viewdirs, const float* v_colors, float* v_coeffs, int N,
                          void* stream) {
    GCB_CHECK_ARG(viewdirs && v_colors && v_coeffs, "null pointer");
    GCB_CHECK_ARG(degree >= 0 && degree <= 3 && K >= (degree + 1) * (degree + 1) && K <= 16, "bad degree=%d / K=%d", degree,
                  K);
    if (N == 0) return GCB_OK;
    const long long total = (long long)N * K;
    long long blocks = (total + 255) / 256;
    if (blocks > 148ll * 16) blocks = 148ll * 16;
    sh_bwd_kernel<<<(unsigned)blocks, 256, 0, ST>>>(degree, K, viewdirs, v_colors, v_coeffs, N);
    GCB_LAUNCH_CHECK();
    return GCB_OK;
}
