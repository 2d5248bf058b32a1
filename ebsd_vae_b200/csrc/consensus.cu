// K3: quaternion misorientation / cubic-symmetry consensus over the top-k hits (float64).
//
// Replaces ChromaLatentVectorDatabase.find_best_orientation / _find_symmetry_equivalent_orientation
// (latice/index/chroma_db.py:261-375) and the FAISS twin (latice/index/faiss_db.py:258-393), which run as a
// serial Python/scipy loop per query.  Here: one warp per query, one lane per candidate.
//
//   candidates  : quaternions gathered from the dictionary's orientation table by the top-k row indices
//   iteration r : ref = candidate r; angle_i = 2 atan2(|xyz|, |w|) of ref*cand_i^-1 (Chroma) or ref^-1*cand_i
//                 (FAISS); similar = {i : angle_i < threshold} in radians (Chroma) or degrees (FAISS);
//                 NO symmetry is applied before thresholding (chroma_db.py:307-310)
//   success     : |similar| >= min_required_matches -> each similar candidate is replaced by its cubic equivalent
//                 closest to ref (first minimum over the 24 operators in table order), then the chordal mean =
//                 dominant eigenvector of sum q q^T (what scipy's Rotation.mean computes), then extrinsic zxz
//                 Euler angles in degrees with scipy's gimbal-lock convention.
#include <math.h>

#include "common.cuh"

namespace ebsd {

struct Quat {
    double x, y, z, w;
};

// 24 proper cubic operators exactly as listed in latice/utils/constants.py:13-38, read scalar-LAST (x,y,z,w)
// the way scipy's Rotation.from_quat reads them.
#define EBSD_S2 0.70710678118654752440
__constant__ double c_cubic[24][4] = {
    {1, 0, 0, 0},
    {0, 1, 0, 0},
    {0, 0, 1, 0},
    {0, 0, 0, 1},
    {0.5, 0.5, 0.5, 0.5},
    {0.5, -0.5, -0.5, -0.5},
    {0.5, 0.5, -0.5, 0.5},
    {0.5, -0.5, 0.5, -0.5},
    {0.5, -0.5, 0.5, 0.5},
    {0.5, 0.5, -0.5, -0.5},
    {0.5, -0.5, -0.5, 0.5},
    {0.5, 0.5, 0.5, -0.5},
    {EBSD_S2, EBSD_S2, 0, 0},
    {EBSD_S2, 0, EBSD_S2, 0},
    {EBSD_S2, 0, 0, EBSD_S2},
    {EBSD_S2, -EBSD_S2, 0, 0},
    {EBSD_S2, 0, -EBSD_S2, 0},
    {EBSD_S2, 0, 0, -EBSD_S2},
    {0, EBSD_S2, EBSD_S2, 0},
    {0, -EBSD_S2, EBSD_S2, 0},
    {0, 0, EBSD_S2, EBSD_S2},
    {0, 0, -EBSD_S2, EBSD_S2},
    {0, EBSD_S2, 0, EBSD_S2},
    {0, -EBSD_S2, 0, EBSD_S2},
};

__device__ __forceinline__ Quat qmul(const Quat &p, const Quat &q) {
    Quat r;
    r.x = p.w * q.x + p.x * q.w + p.y * q.z - p.z * q.y;
    r.y = p.w * q.y - p.x * q.z + p.y * q.w + p.z * q.x;
    r.z = p.w * q.z + p.x * q.y - p.y * q.x + p.z * q.w;
    r.w = p.w * q.w - p.x * q.x - p.y * q.y - p.z * q.z;
    return r;
}
__device__ __forceinline__ Quat qconj(const Quat &q) { return Quat{-q.x, -q.y, -q.z, q.w}; }
__device__ __forceinline__ double qangle(const Quat &q) {
    return 2.0 * atan2(sqrt(q.x * q.x + q.y * q.y + q.z * q.z), fabs(q.w));
}
__device__ __forceinline__ double shfl_d(double v, int src) { return __shfl_sync(0xffffffffu, v, src); }
__device__ __forceinline__ Quat shfl_q(const Quat &q, int src) {
    return Quat{shfl_d(q.x, src), shfl_d(q.y, src), shfl_d(q.z, src), shfl_d(q.w, src)};
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Dominant eigenvector of a symmetric 4x4 matrix by cyclic Jacobi rotations.
__device__ void dominant_eigvec4(double a[4][4], double out[4]) {
    double v[4][4] = {{1, 0, 0, 0}, {0, 1, 0, 0}, {0, 0, 1, 0}, {0, 0, 0, 1}};
    for (int sweep = 0; sweep < 32; ++sweep) {
        double off = 0.0;
        for (int p = 0; p < 4; ++p)
            for (int q = p + 1; q < 4; ++q) off += a[p][q] * a[p][q];
        if (off < 1e-300) break;
        for (int p = 0; p < 4; ++p) {
            for (int q = p + 1; q < 4; ++q) {
                const double apq = a[p][q];
                if (fabs(apq) < 1e-300) continue;
                const double theta = (a[q][q] - a[p][p]) / (2.0 * apq);
                const double t = (theta >= 0.0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
                const double c = 1.0 / sqrt(t * t + 1.0), s = t * c;
                for (int r = 0; r < 4; ++r) {  // A <- A J
                    const double arp = a[r][p], arq = a[r][q];
                    a[r][p] = c * arp - s * arq;
                    a[r][q] = s * arp + c * arq;
                }
                for (int r = 0; r < 4; ++r) {  // A <- J^T A
                    const double apr = a[p][r], aqr = a[q][r];
                    a[p][r] = c * apr - s * aqr;
                    a[q][r] = s * apr + c * aqr;
                }
                for (int r = 0; r < 4; ++r) {
                    const double vrp = v[r][p], vrq = v[r][q];
                    v[r][p] = c * vrp - s * vrq;
                    v[r][q] = s * vrp + c * vrq;
                }
            }
        }
    }
    int best = 0;
    for (int i = 1; i < 4; ++i)
        if (a[i][i] > a[best][best]) best = i;
    double n = 0.0;
    for (int r = 0; r < 4; ++r) n += v[r][best] * v[r][best];
    n = 1.0 / sqrt(n);
    for (int r = 0; r < 4; ++r) out[r] = v[r][best] * n;
}

// scipy Rotation.as_euler("zxz", degrees=True) for a unit quaternion (x,y,z,w).
__device__ void euler_zxz_deg(const Quat &q, double out[3]) {
    const double kPi = 3.14159265358979323846;
    const double half_sum = atan2(q.z, q.w);
    const double half_diff = atan2(-q.y, q.x);
    const double big = 2.0 * atan2(hypot(q.x, q.y), hypot(q.w, q.z));
    double first, third;
    if (fabs(big) <= 1e-7) {
        first = 2.0 * half_sum;
        third = 0.0;
    } else if (fabs(big - kPi) <= 1e-7) {
        first = 2.0 * half_diff;
        third = 0.0;
    } else {
        first = half_sum + half_diff;
        third = half_sum - half_diff;
    }
    double ang[3] = {first, big, third};
    for (int i = 0; i < 3; ++i) {
        if (ang[i] < -kPi) ang[i] += 2.0 * kPi;
        else if (ang[i] > kPi) ang[i] -= 2.0 * kPi;
        out[i] = ang[i] * (180.0 / kPi);
    }
}

struct ConsensusParams {
    const double *quat_table;
    const double *euler_table;   // nullable: [N,3] degrees as stored, gathered into cand_euler
    double *cand_euler;          // nullable: [Q,k,3], NaN where the candidate slot is empty
    long long N;
    long long index_base;        // global row of table row 0 (cand_idx holds global rows)
    const long long *cand_idx;
    long long Q;
    int k;
    double threshold;
    int degrees;
    int min_required;
    int max_iter;
    int faiss;
    double *mean_quat;
    double *mean_euler;
    uint8_t *success;
    unsigned long long *similar_mask;
    int *ref_iter;
};

// Resident blocks per SM: the kernel is a chain of fp64 latencies (divisions, square roots, acos / atan2) with one warp
// per query, so its speed is the number of resident warps.  Uncapped it needs 132 registers = ONE block of 8 warps per SM.
#ifndef EBSD_CONSENSUS_MINBLOCKS
#define EBSD_CONSENSUS_MINBLOCKS 2
#endif
__global__ void __launch_bounds__(256, EBSD_CONSENSUS_MINBLOCKS) consensus_kernel(const ConsensusParams p) {
    const long long q = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (q >= p.Q) return;
    const double kNaN = __longlong_as_double(0x7ff8000000000000ll);

    long long row = -1;
    if (lane < p.k) {
        row = p.cand_idx[q * p.k + lane];
        if (row >= 0) row -= p.index_base;
    }
    const bool have = row >= 0 && row < p.N;
    if (p.cand_euler && lane < p.k) {
        // candidate_orientations of OrientationResult (chroma_db.py:283-289): the stored Euler triplets, bit for bit
        double e0 = kNaN, e1 = kNaN, e2 = kNaN;
        if (have && p.euler_table) {
            e0 = p.euler_table[row * 3 + 0];
            e1 = p.euler_table[row * 3 + 1];
            e2 = p.euler_table[row * 3 + 2];
        }
        double *dst = p.cand_euler + (q * p.k + lane) * 3;
        dst[0] = e0;
        dst[1] = e1;
        dst[2] = e2;
    }
    Quat mine = {0, 0, 0, 1};
    if (have) {
        const double4 t = *(const double4 *)(p.quat_table + row * 4);
        mine = Quat{t.x, t.y, t.z, t.w};
    }
    const unsigned have_mask = __ballot_sync(0xffffffffu, have);
    const int k_valid = __popc(have_mask);
    const int iters = p.max_iter < k_valid ? p.max_iter : k_valid;

    bool ok = false;
    unsigned sim_mask = 0;
    int ref_it = -1;
    Quat ref = {0, 0, 0, 1};
    for (int it = 0; it < iters; ++it) {
        ref = shfl_q(mine, it);
        ref_it = it;
        const Quat rel = p.faiss ? qmul(qconj(ref), mine) : qmul(ref, qconj(mine));
        double ang = qangle(rel);
        if (p.degrees) ang = ang * (180.0 / 3.14159265358979323846);
        sim_mask = __ballot_sync(0xffffffffu, have && (ang < p.threshold));
        if (__popc(sim_mask) >= p.min_required) {
            ok = true;
            break;
        }
    }

    double mq[4] = {kNaN, kNaN, kNaN, kNaN};
    double me[3] = {kNaN, kNaN, kNaN};
    if (ok && sim_mask != 0) {
        const bool similar = (sim_mask >> lane) & 1u;
        Quat red = {0, 0, 0, 0};
        if (similar) {
            double best = 1e300;
            int best_j = 0;
            if (!p.faiss) {
                const Quat cinv = qconj(mine);
                for (int j = 0; j < 24; ++j) {
                    const Quat s = Quat{c_cubic[j][0], c_cubic[j][1], c_cubic[j][2], c_cubic[j][3]};
                    const double a = qangle(qmul(ref, qmul(cinv, s)));
                    if (a < best) {
                        best = a;
                        best_j = j;
                    }
                }
                const Quat s = Quat{c_cubic[best_j][0], c_cubic[best_j][1], c_cubic[best_j][2], c_cubic[best_j][3]};
                red = qconj(qmul(cinv, s));  // (cand^-1 S)^-1 = S^-1 cand
            } else {
                const Quat rinv = qconj(ref);
                for (int j = 0; j < 24; ++j) {
                    const Quat s = Quat{c_cubic[j][0], c_cubic[j][1], c_cubic[j][2], c_cubic[j][3]};
                    const double a = qangle(qmul(rinv, qmul(s, mine)));
                    if (a < best) {
                        best = a;
                        best_j = j;
                    }
                }
                const Quat s = Quat{c_cubic[best_j][0], c_cubic[best_j][1], c_cubic[best_j][2], c_cubic[best_j][3]};
                red = qmul(s, mine);
            }
        }
        const double c[4] = {red.x, red.y, red.z, red.w};
        double a[4][4];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = i; j < 4; ++j) {
                const double s = warp_sum(c[i] * c[j]);
                a[i][j] = s;
                a[j][i] = s;
            }
        if (lane == 0) {
            dominant_eigvec4(a, mq);
            if (mq[3] < 0.0) {
                for (int i = 0; i < 4; ++i) mq[i] = -mq[i];
            }
            euler_zxz_deg(Quat{mq[0], mq[1], mq[2], mq[3]}, me);
        }
    }
    if (lane == 0) {
        for (int i = 0; i < 4; ++i) p.mean_quat[q * 4 + i] = mq[i];
        for (int i = 0; i < 3; ++i) p.mean_euler[q * 3 + i] = me[i];
        p.success[q] = ok ? 1 : 0;
        p.similar_mask[q] = (unsigned long long)sim_mask;
        p.ref_iter[q] = ref_it;
    }
}

// scipy R.from_euler("zxz", [a, b, c], degrees=True) as a quaternion (x, y, z, w): extrinsic, R = Rz(c) Rx(b) Rz(a)
__device__ __forceinline__ Quat euler_zxz_to_quat(double a_deg, double b_deg, double c_deg) {
    const double kRad = 3.14159265358979323846 / 180.0;
    const double a = a_deg * kRad, b = b_deg * kRad, c = c_deg * kRad;
    double sb, cb, sp, cp, sm, cm;
    sincos(0.5 * b, &sb, &cb);
    sincos(0.5 * (a + c), &sp, &cp);
    sincos(0.5 * (a - c), &sm, &cm);
    return Quat{sb * cm, -sb * sm, cb * sp, cb * cp};
}

__global__ void euler_to_quat_kernel(const double *euler_deg, long long n, double *quat) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const Quat q = euler_zxz_to_quat(euler_deg[i * 3 + 0], euler_deg[i * 3 + 1], euler_deg[i * 3 + 2]);
    *(double4 *)(quat + i * 4) = make_double4(q.x, q.y, q.z, q.w);
}

// ---------------------------------------------------------------------------------------------------------------
// IPF colour key: get_color_key (latice/utils/utils.py:206-240) +
// ColorKeyGenerator.generate_ipf_color (latice/utils/colorkey.py:64-130).  One thread per orientation, float64,
// products and sums un-fused in the reference's order so that the unit-triangle decisions agree.
// ---------------------------------------------------------------------------------------------------------------
// QUAT_SYM.as_matrix() (scipy, row-major 3x3 per operator) for the table above, to 17 significant digits
__constant__ double c_cubic_mat[24][9] = {
    {1.0, 0.0, 0.0, 0.0, -1.0, 0.0, 0.0, 0.0, -1.0},
    {-1.0, 0.0, 0.0, 0.0, 1.0, 0.0, 0.0, 0.0, -1.0},
    {-1.0, 0.0, 0.0, 0.0, -1.0, 0.0, 0.0, 0.0, 1.0},
    {1.0, 0.0, 0.0, 0.0, 1.0, 0.0, 0.0, 0.0, 1.0},
    {0.0, 0.0, 1.0, 1.0, 0.0, 0.0, 0.0, 1.0, 0.0},
    {0.0, -1.0, 0.0, 0.0, 0.0, 1.0, -1.0, 0.0, 0.0},
    {0.0, 1.0, 0.0, 0.0, 0.0, -1.0, -1.0, 0.0, 0.0},
    {0.0, 0.0, 1.0, -1.0, 0.0, 0.0, 0.0, -1.0, 0.0},
    {0.0, -1.0, 0.0, 0.0, 0.0, -1.0, 1.0, 0.0, 0.0},
    {0.0, 0.0, -1.0, 1.0, 0.0, 0.0, 0.0, -1.0, 0.0},
    {0.0, 0.0, -1.0, -1.0, 0.0, 0.0, 0.0, 1.0, 0.0},
    {0.0, 1.0, 0.0, 0.0, 0.0, 1.0, 1.0, 0.0, 0.0},
    {0.0, 1.0000000000000002, 0.0, 1.0000000000000002, 0.0, 0.0, 0.0, 0.0, -1.0000000000000002},
    {0.0, 0.0, 1.0000000000000002, 0.0, -1.0000000000000002, 0.0, 1.0000000000000002, 0.0, 0.0},
    {1.0000000000000002, 0.0, 0.0, 0.0, 0.0, -1.0000000000000002, 0.0, 1.0000000000000002, 0.0},
    {0.0, -1.0000000000000002, 0.0, -1.0000000000000002, 0.0, -0.0, 0.0, 0.0, -1.0000000000000002},
    {0.0, 0.0, -1.0000000000000002, 0.0, -1.0000000000000002, -0.0, -1.0000000000000002, 0.0, 0.0},
    {1.0000000000000002, 0.0, 0.0, 0.0, 0.0, 1.0000000000000002, 0.0, -1.0000000000000002, 0.0},
    {-1.0000000000000002, 0.0, 0.0, 0.0, 0.0, 1.0000000000000002, 0.0, 1.0000000000000002, 0.0},
    {-1.0000000000000002, -0.0, 0.0, 0.0, 0.0, -1.0000000000000002, 0.0, -1.0000000000000002, 0.0},
    {0.0, -1.0000000000000002, 0.0, 1.0000000000000002, 0.0, 0.0, 0.0, 0.0, 1.0000000000000002},
    {0.0, 1.0000000000000002, 0.0, -1.0000000000000002, 0.0, -0.0, -0.0, 0.0, 1.0000000000000002},
    {0.0, 0.0, 1.0000000000000002, 0.0, 1.0000000000000002, 0.0, -1.0000000000000002, 0.0, 0.0},
    {0.0, -0.0, -1.0000000000000002, 0.0, 1.0000000000000002, -0.0, 1.0000000000000002, 0.0, 0.0},
};

__device__ __forceinline__ double dot3_unfused(const double *m, double x, double y, double z) {
    return __dadd_rn(__dadd_rn(__dmul_rn(m[0], x), __dmul_rn(m[1], y)), __dmul_rn(m[2], z));
}

__global__ void ipf_color_kernel(const double *__restrict__ euler_deg, long long n, int row, uint8_t *__restrict__ rgb) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const Quat q = euler_zxz_to_quat(euler_deg[i * 3 + 0], euler_deg[i * 3 + 1], euler_deg[i * 3 + 2]);
    // scipy as_matrix of a unit quaternion, row `row`
    const double x2 = __dmul_rn(q.x, q.x), y2 = __dmul_rn(q.y, q.y), z2 = __dmul_rn(q.z, q.z), w2 = __dmul_rn(q.w, q.w);
    const double xy = __dmul_rn(q.x, q.y), zw = __dmul_rn(q.z, q.w), xz = __dmul_rn(q.x, q.z), yw = __dmul_rn(q.y, q.w);
    const double yz = __dmul_rn(q.y, q.z), xw = __dmul_rn(q.x, q.w);
    double px, py, pz;
    if (row == 0) {
        px = __dadd_rn(__dsub_rn(__dsub_rn(x2, y2), z2), w2);
        py = __dmul_rn(2.0, __dsub_rn(xy, zw));
        pz = __dmul_rn(2.0, __dadd_rn(xz, yw));
    } else if (row == 1) {
        px = __dmul_rn(2.0, __dadd_rn(xy, zw));
        py = __dadd_rn(__dsub_rn(__dadd_rn(-x2, y2), z2), w2);
        pz = __dmul_rn(2.0, __dsub_rn(yz, xw));
    } else {
        px = __dmul_rn(2.0, __dsub_rn(xz, yw));
        py = __dmul_rn(2.0, __dadd_rn(yz, xw));
        pz = __dadd_rn(__dadd_rn(__dsub_rn(-x2, y2), z2), w2);
    }
    const double nrm = sqrt(__dadd_rn(__dadd_rn(__dmul_rn(px, px), __dmul_rn(py, py)), __dmul_rn(pz, pz)));
    px = px / nrm;
    py = py / nrm;
    pz = pz / nrm;
    const double eta_max = 45.0 * (3.14159265358979323846 / 180.0);
    const double chi_lim = acos(1.0 / sqrt(3.0));
    double chi = 0.0, eta = 0.0;
    for (int c = 0; c < 48; ++c) {
        const double *m = c_cubic_mat[c % 24];
        double vx = dot3_unfused(m, px, py, pz), vy = dot3_unfused(m + 3, px, py, pz), vz = dot3_unfused(m + 6, px, py, pz);
        if (c >= 24) {
            vx = -vx;
            vy = -vy;
            vz = -vz;
        }
        if (vz < 0) {  // USE_INVERSION (latice/utils/constants.py:11)
            vx = -vx;
            vy = -vy;
            vz = -vz;
        }
        chi = acos(fmin(1.0, fmax(-1.0, vz)));
        eta = atan2(vy, vx);
        if (!(eta < 0 || eta > eta_max || chi < 0 || chi > chi_lim)) break;
    }
    const double k = 180.0 / 3.14159265358979323846;
    const double chi_max = chi_lim * k;
    const double eta_deg = eta * k, chi_deg = chi * k;
    double r = 1.0 - chi_deg / chi_max, b = fabs(eta_deg - 0.0) / 45.0;
    double g = 1.0 - b;
    g = g * (chi_deg / chi_max);
    b = b * (chi_deg / chi_max);
    r = sqrt(r);
    g = sqrt(g);
    b = sqrt(b);
    const double mx = fmax(r, fmax(g, b));
    // Python round(): half to even, exactly what rint does in the default rounding mode
    rgb[i * 3 + 0] = (uint8_t)(int)rint(255.0 * r / mx);
    rgb[i * 3 + 1] = (uint8_t)(int)rint(255.0 * g / mx);
    rgb[i * 3 + 2] = (uint8_t)(int)rint(255.0 * b / mx);
}

}  // namespace ebsd

using namespace ebsd;

extern "C" {

int ebsd_euler_to_quat(const double *euler_deg, int64_t n, double *quat, void *stream) {
    int rc = check_device_arch();
    if (rc) return rc;
    EBSD_REQUIRE(n >= 0, "ebsd_euler_to_quat: negative n");
    if (n == 0) return EBSD_OK;
    EBSD_REQUIRE(euler_deg && quat, "ebsd_euler_to_quat: null pointer");
    EBSD_REQUIRE(((uintptr_t)quat & 31) == 0, "ebsd_euler_to_quat: quat must be 32-byte aligned");
    const int threads = 256;
    euler_to_quat_kernel<<<(unsigned)((n + threads - 1) / threads), threads, 0, (cudaStream_t)stream>>>(euler_deg, n,
                                                                                                       quat);
    EBSD_LAUNCH_CHECK();
    return EBSD_OK;
}

int ebsd_ipf_color(const double *euler_deg, int64_t n, int axis, uint8_t *rgb, void *stream) {
    int rc = check_device_arch();
    if (rc) return rc;
    EBSD_REQUIRE(n >= 0, "ebsd_ipf_color: negative n");
    EBSD_REQUIRE(axis >= 0 && axis <= 2, "ebsd_ipf_color: axis must be 0 (ipf_x), 1 (ipf_y) or 2 (ipf_z), got %d", axis);
    if (n == 0) return EBSD_OK;
    EBSD_REQUIRE(euler_deg && rgb, "ebsd_ipf_color: null pointer");
    const int threads = 128;
    ipf_color_kernel<<<(unsigned)((n + threads - 1) / threads), threads, 0, (cudaStream_t)stream>>>(euler_deg, n, axis,
                                                                                                   rgb);
    EBSD_LAUNCH_CHECK();
    return EBSD_OK;
}

int ebsd_consensus(const double *quat_table, const double *euler_table, int64_t N, int64_t index_base,
                   const int64_t *cand_idx, int64_t Q, int k, double threshold, int angle_unit,
                   int min_required_matches, int max_iterations, int faiss_semantics, double *mean_quat,
                   double *mean_euler_deg, uint8_t *success, uint64_t *similar_mask, int32_t *ref_iter,
                   double *cand_euler_deg, void *stream) {
    int rc = check_device_arch();
    if (rc) return rc;
    EBSD_REQUIRE(k >= 1 && k <= EBSD_MAX_TOPK, "ebsd_consensus: k must be in [1,%d], got %d", EBSD_MAX_TOPK, k);
    EBSD_REQUIRE(Q >= 0 && N >= 0, "ebsd_consensus: negative size");
    EBSD_REQUIRE(angle_unit == EBSD_ANGLE_RADIANS || angle_unit == EBSD_ANGLE_DEGREES,
                 "ebsd_consensus: bad angle_unit %d", angle_unit);
    if (Q == 0) return EBSD_OK;
    EBSD_REQUIRE(cand_idx && mean_quat && mean_euler_deg && success && similar_mask && ref_iter,
                 "ebsd_consensus: null pointer");
    EBSD_REQUIRE(N == 0 || quat_table != nullptr, "ebsd_consensus: null orientation table");
    EBSD_REQUIRE(((uintptr_t)quat_table & 31) == 0, "ebsd_consensus: quat_table must be 32-byte aligned");
    ConsensusParams p;
    p.quat_table = quat_table;
    p.euler_table = euler_table;
    p.cand_euler = cand_euler_deg;
    p.N = N;
    p.index_base = index_base;
    p.cand_idx = (const long long *)cand_idx;
    p.Q = Q;
    p.k = k;
    p.threshold = threshold;
    p.degrees = angle_unit == EBSD_ANGLE_DEGREES;
    p.min_required = min_required_matches;
    p.max_iter = max_iterations;
    p.faiss = faiss_semantics != 0;
    p.mean_quat = mean_quat;
    p.mean_euler = mean_euler_deg;
    p.success = success;
    p.similar_mask = (unsigned long long *)similar_mask;
    p.ref_iter = ref_iter;
    const int wpb = 8;
    consensus_kernel<<<(unsigned)((Q + wpb - 1) / wpb), wpb * 32, 0, (cudaStream_t)stream>>>(p);
    EBSD_LAUNCH_CHECK();
    return EBSD_OK;
}

}  // extern "C"
