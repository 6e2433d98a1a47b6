// SMPL / SMPLH linear blend skinning on device: the step that feeds the rasteriser every frame.
// Mirrors iPERCore/tools/human_digitalizer/smplx/lbs.py:137-227 (lbs), :321-375 (batch_rigid_transform),
// iPERCore/tools/utils/geometry/rotations.py:318-332,355-375 (axis-angle -> quaternion -> matrix) and
// bodynets/base_smpl.py:28-50 (link).  Shape (betas, offsets) is constant per source in run_imitator
// (imitator.py:248-256 swaps in the SOURCE shape), so the shape-dependent part (v_shaped, rest joints) is a one-time
// setup pair of kernels and the per-frame work is: 52 Rodrigues + a 52-link kinematic chain per frame (one small CTA per
// frame), then one pass over the vertices that applies the 459-term pose blend shapes and the 52-joint skinning for a
// chunk of frames at a time (posedirs, 38 MB, stays L2-resident across chunks).
#include "common.cuh"
#include "iper_b200.h"

namespace iper {

constexpr int LBS_MAX_J = 64;
constexpr int LBS_FB = 8;        // frames per vertex-kernel CTA

// ---- one-time per shape ----------------------------------------------------------------------------------------
// v_shaped[v,k] = v_template[v,k] + offsets[v,k] + sum_l betas[l] * shapedirs[v,k,l]        (lbs.py:182, :259-271)
__global__ void lbs_shape_kernel(const float* __restrict__ v_template, const float* __restrict__ offsets,
                                 const float* __restrict__ shapedirs, const float* __restrict__ betas, int nv, int nb,
                                 float* __restrict__ v_shaped) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;      // over nv*3
    if (i >= nv * 3) return;
    float acc = 0.f;
    for (int l = 0; l < nb; l++) acc += betas[l] * shapedirs[(size_t)i * nb + l];
    v_shaped[i] = (v_template[i] + (offsets ? offsets[i] : 0.f)) + acc;
}
// J[j,k] = sum_v J_regressor[j,v] * v_shaped[v,k]                                            (lbs.py:187, vertices2joints)
__global__ void lbs_joints_rest_kernel(const float* __restrict__ J_regressor, const float* __restrict__ v_shaped, int nv,
                                       float* __restrict__ J) {
    const int j = blockIdx.x;
    float a[3] = {0.f, 0.f, 0.f};
    for (int v = threadIdx.x; v < nv; v += blockDim.x) {
        const float w = J_regressor[(size_t)j * nv + v];
        if (w != 0.f) { a[0] += w * v_shaped[3 * v]; a[1] += w * v_shaped[3 * v + 1]; a[2] += w * v_shaped[3 * v + 2]; }
    }
    __shared__ float red[3][256];
    for (int k = 0; k < 3; k++) red[k][threadIdx.x] = a[k];
    __syncthreads();
    for (int s = blockDim.x / 2; s > 0; s >>= 1) {
        if (threadIdx.x < s) for (int k = 0; k < 3; k++) red[k][threadIdx.x] += red[k][threadIdx.x + s];
        __syncthreads();
    }
    if (threadIdx.x < 3) J[3 * j + threadIdx.x] = red[threadIdx.x][0];
}

// ---- per frame: rotations, pose feature, kinematic chain -----------------------------------------------------------
// pose (B, nj*3) axis-angle -> pose_feature (B, (nj-1)*9) = R[1:] - I ; A (B, nj, 12) = rows 0..2 of the relative rigid
// transforms (lbs.py:372-373) ; posed joints (B, nj, 3)
__global__ void lbs_pose_kernel(const float* __restrict__ pose, const float* __restrict__ J, const int* __restrict__ parents,
                                int nj, float* __restrict__ pose_feature, float* __restrict__ A,
                                float* __restrict__ joints_out) {
    __shared__ float R[LBS_MAX_J][9];
    __shared__ float G[LBS_MAX_J][12];      // world transform rows 0..2 of every joint
    const int b = blockIdx.x, j = threadIdx.x;
    if (j < nj) {
        const float* p = pose + ((size_t)b * nj + j) * 3;
        // rotvec_to_rotmat: angle = ||v + 1e-8||, q = (cos(a/2), sin(a/2) * v/angle), normalised, then the quaternion formula
        const float x0 = p[0], y0 = p[1], z0 = p[2];
        const float ex = x0 + 1e-8f, ey = y0 + 1e-8f, ez = z0 + 1e-8f;
        const float angle = sqrtf(ex * ex + ey * ey + ez * ez);
        const float half = angle * 0.5f, sn = sinf(half), cs = cosf(half);
        float w = cs, x = sn * (x0 / angle), y = sn * (y0 / angle), z = sn * (z0 / angle);
        const float qn = sqrtf(w * w + x * x + y * y + z * z);
        w /= qn; x /= qn; y /= qn; z /= qn;
        const float w2 = w * w, x2 = x * x, y2 = y * y, z2 = z * z;
        const float wx = w * x, wy = w * y, wz = w * z, xy = x * y, xz = x * z, yz = y * z;
        float r[9] = {w2 + x2 - y2 - z2, 2 * xy - 2 * wz, 2 * wy + 2 * xz,
                      2 * wz + 2 * xy, w2 - x2 + y2 - z2, 2 * yz - 2 * wx,
                      2 * xz - 2 * wy, 2 * wx + 2 * yz, w2 - x2 - y2 + z2};
        for (int k = 0; k < 9; k++) R[j][k] = r[k];
        if (j > 0) {
            float* pf = pose_feature + ((size_t)b * (nj - 1) + (j - 1)) * 9;
            for (int k = 0; k < 9; k++) pf[k] = r[k] - ((k == 0 || k == 4 || k == 8) ? 1.f : 0.f);
        }
    }
    __syncthreads();
    if (j == 0) {       // 52-link chain: G_i = G_parent(i) * [R_i | J_i - J_parent(i)]   (lbs.py:345-358)
        for (int k = 0; k < 9; k++) G[0][(k / 3) * 4 + k % 3] = R[0][k];
        for (int k = 0; k < 3; k++) G[0][k * 4 + 3] = J[k];
        for (int i = 1; i < nj; i++) {
            const int pa = parents[i];
            float t[3];
            for (int k = 0; k < 3; k++) t[k] = J[3 * i + k] - J[3 * pa + k];
            for (int r = 0; r < 3; r++) {
                const float g0 = G[pa][r * 4], g1 = G[pa][r * 4 + 1], g2 = G[pa][r * 4 + 2], g3 = G[pa][r * 4 + 3];
                for (int c = 0; c < 3; c++) G[i][r * 4 + c] = g0 * R[i][c] + g1 * R[i][3 + c] + g2 * R[i][6 + c];
                G[i][r * 4 + 3] = g0 * t[0] + g1 * t[1] + g2 * t[2] + g3;
            }
        }
    }
    __syncthreads();
    if (j < nj) {
        float* a = A + ((size_t)b * nj + j) * 12;
        for (int r = 0; r < 3; r++) {
            const float g0 = G[j][r * 4], g1 = G[j][r * 4 + 1], g2 = G[j][r * 4 + 2];
            a[r * 4] = g0; a[r * 4 + 1] = g1; a[r * 4 + 2] = g2;
            a[r * 4 + 3] = G[j][r * 4 + 3] - (g0 * J[3 * j] + g1 * J[3 * j + 1] + g2 * J[3 * j + 2]);
            if (joints_out) joints_out[((size_t)b * nj + j) * 3 + r] = G[j][r * 4 + 3];
        }
    }
}

// ---- per vertex: pose blend shapes + skinning for LBS_FB frames --------------------------------------------------------
// posedirs (P, nv*3) row-major ; weights_t (nj, nv) = lbs_weights transposed ; src_of (nv) = vertex whose skinned position
// output vertex v takes (identity, or the cloth link source: base_smpl.py:41-44)
__global__ void __launch_bounds__(128) lbs_skin_kernel(const float* __restrict__ v_shaped, const float* __restrict__ posedirs,
                                                       const float* __restrict__ weights_t, const int* __restrict__ src_of,
                                                       const float* __restrict__ pose_feature, const float* __restrict__ A,
                                                       int B, int nv, int nj, float* __restrict__ verts) {
    extern __shared__ float sm[];
    const int P = (nj - 1) * 9;
    float* s_pf = sm;                      // [LBS_FB][P]
    float* s_A = sm + LBS_FB * P;          // [LBS_FB][nj][12]
    const int b0 = blockIdx.y * LBS_FB, nb = min(LBS_FB, B - b0);
    for (int i = threadIdx.x; i < nb * P; i += blockDim.x) s_pf[i] = pose_feature[(size_t)b0 * P + i];
    for (int i = threadIdx.x; i < nb * nj * 12; i += blockDim.x) s_A[i] = A[(size_t)b0 * nj * 12 + i];
    __syncthreads();
    const int vo = blockIdx.x * blockDim.x + threadIdx.x;
    if (vo >= nv) return;
    const int v = src_of ? src_of[vo] : vo;
    float acc[LBS_FB][3];
#pragma unroll
    for (int f = 0; f < LBS_FB; f++) { acc[f][0] = 0.f; acc[f][1] = 0.f; acc[f][2] = 0.f; }
    for (int p = 0; p < P; p++) {          // pose_offsets = pose_feature @ posedirs   (lbs.py:204)
        const float* pd = posedirs + (size_t)p * nv * 3 + 3 * v;
        const float d0 = __ldg(pd), d1 = __ldg(pd + 1), d2 = __ldg(pd + 2);
#pragma unroll
        for (int f = 0; f < LBS_FB; f++) {
            const float w = s_pf[f * P + p];
            acc[f][0] += w * d0; acc[f][1] += w * d1; acc[f][2] += w * d2;
        }
    }
    const float vs0 = v_shaped[3 * v], vs1 = v_shaped[3 * v + 1], vs2 = v_shaped[3 * v + 2];
    float T[LBS_FB][12];
#pragma unroll
    for (int f = 0; f < LBS_FB; f++)
#pragma unroll
        for (int k = 0; k < 12; k++) T[f][k] = 0.f;
    for (int j = 0; j < nj; j++) {         // T = W @ A   (lbs.py:218)
        const float w = __ldg(weights_t + (size_t)j * nv + v);
        if (w == 0.f) continue;
#pragma unroll
        for (int f = 0; f < LBS_FB; f++)
#pragma unroll
            for (int k = 0; k < 12; k++) T[f][k] += w * s_A[(f * nj + j) * 12 + k];
    }
#pragma unroll
    for (int f = 0; f < LBS_FB; f++) {
        if (f >= nb) break;
        const float p0 = acc[f][0] + vs0, p1 = acc[f][1] + vs1, p2 = acc[f][2] + vs2;   // v_posed = pose_offsets + v_shaped
        float* o = verts + ((size_t)(b0 + f) * nv + vo) * 3;
        o[0] = T[f][0] * p0 + T[f][1] * p1 + T[f][2] * p2 + T[f][3];
        o[1] = T[f][4] * p0 + T[f][5] * p1 + T[f][6] * p2 + T[f][7];
        o[2] = T[f][8] * p0 + T[f][9] * p1 + T[f][10] * p2 + T[f][11];
    }
}

}  // namespace iper

using namespace iper;

extern "C" int iper_lbs_shape(const float* v_template, const float* offsets, const float* shapedirs, const float* betas,
                              const float* J_regressor, int nv, int nb, int nj, float* v_shaped, float* J_rest,
                              iper_stream_t stream) {
    IPER_REQUIRE(v_template && shapedirs && betas && J_regressor && v_shaped && J_rest, "iper_lbs_shape: null pointer");
    IPER_REQUIRE(nj >= 1 && nj <= LBS_MAX_J, "iper_lbs_shape: nj=%d not in [1,%d]", nj, LBS_MAX_J);
    cudaStream_t s = (cudaStream_t)stream;
    lbs_shape_kernel<<<(nv * 3 + 255) / 256, 256, 0, s>>>(v_template, offsets, shapedirs, betas, nv, nb, v_shaped);
    IPER_CHECK_CUDA(cudaGetLastError());
    lbs_joints_rest_kernel<<<nj, 256, 0, s>>>(J_regressor, v_shaped, nv, J_rest);
    IPER_CHECK_CUDA(cudaGetLastError());
    return 0;
}

extern "C" int iper_lbs_frames(const float* pose, int B, int nj, const float* v_shaped, const float* J_rest,
                               const int32_t* parents, const float* posedirs, const float* weights_t,
                               const int32_t* src_of, int nv, float* pose_feature, float* A, float* joints,
                               float* verts, iper_stream_t stream) {
    IPER_REQUIRE(pose && v_shaped && J_rest && parents && posedirs && weights_t && pose_feature && A && verts,
                 "iper_lbs_frames: null pointer");
    IPER_REQUIRE(nj >= 1 && nj <= LBS_MAX_J, "iper_lbs_frames: nj=%d not in [1,%d]", nj, LBS_MAX_J);
    if (B == 0) return 0;
    cudaStream_t s = (cudaStream_t)stream;
    lbs_pose_kernel<<<B, LBS_MAX_J, 0, s>>>(pose, J_rest, parents, nj, pose_feature, A, joints);
    IPER_CHECK_CUDA(cudaGetLastError());
    const size_t smem = sizeof(float) * ((size_t)LBS_FB * (nj - 1) * 9 + (size_t)LBS_FB * nj * 12);
    IPER_CHECK_CUDA(cudaFuncSetAttribute(lbs_skin_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dim3 grid((nv + 127) / 128, (B + LBS_FB - 1) / LBS_FB);
    lbs_skin_kernel<<<grid, 128, smem, s>>>(v_shaped, posedirs, weights_t, src_of, pose_feature, A, B, nv, nj, verts);
    IPER_CHECK_CUDA(cudaGetLastError());
    return 0;
}
