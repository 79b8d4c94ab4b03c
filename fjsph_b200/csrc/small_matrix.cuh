// small_matrix.cuh — per-particle 3x3 helpers run once per particle (never per pair).
// These mirror what FJSPH gets from Eigen in dSPH_PreStep (reference src/Shifting.cpp:73-99):
//   ColPivHouseholderQR<3x3>::isInvertible / inverse   and
//   SelfAdjointEigenSolver<3x3>::computeDirect(...).eigenvalues().minCoeff()
#pragma once
#include <cuda_runtime.h>
#include <float.h>

// Householder QR with column pivoting; rank test |R_kk| > eps*3*max|R_kk|; inverse by solving A x = e_c.
// a, inv: row-major [3][3].  Returns 1 when invertible (inv filled), 0 otherwise (inv untouched).
__device__ __forceinline__ int fj_qr_inverse3(const double (&a)[3][3], double (&inv)[3][3])
{
    double qr[3][3];
    double hco[3];
    int perm[3] = {0, 1, 2};
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) qr[i][j] = a[i][j];

    double maxcol = 0.0;
#pragma unroll
    for (int j = 0; j < 3; ++j)
    {
        double s = qr[0][j] * qr[0][j] + qr[1][j] * qr[1][j] + qr[2][j] * qr[2][j];
        maxcol = fmax(maxcol, sqrt(s));
    }
    const double th = (maxcol * DBL_EPSILON / 3.0);
    const double thr_helper = th * th;
    int nonzero_pivots = 3;
    double maxpivot = 0.0;
#pragma unroll
    for (int k = 0; k < 3; ++k)
    {
        int big = k;
        double bigsq = -1.0;
#pragma unroll
        for (int j = 0; j < 3; ++j)
        {
            if (j >= k)
            {
                double s = 0.0;
#pragma unroll
                for (int i = 0; i < 3; ++i)
                    if (i >= k)
                        s += qr[i][j] * qr[i][j];
                if (s > bigsq)
                {
                    bigsq = s;
                    big = j;
                }
            }
        }
        if (nonzero_pivots == 3 && bigsq < thr_helper * double(3 - k))
            nonzero_pivots = k;
        if (big != k)
        {
#pragma unroll
            for (int j = 0; j < 3; ++j)
            {
                if (j == big) // static indexing only: swap column k with column j
                {
#pragma unroll
                    for (int i = 0; i < 3; ++i)
                    {
                        double t = qr[i][k];
                        qr[i][k] = qr[i][j];
                        qr[i][j] = t;
                    }
                    int tp = perm[k];
                    perm[k] = perm[j];
                    perm[j] = tp;
                }
            }
        }
        double c0 = qr[k][k];
        double tailsq = 0.0;
#pragma unroll
        for (int i = 0; i < 3; ++i)
            if (i > k)
                tailsq += qr[i][k] * qr[i][k];
        double beta, tau;
        if (tailsq <= DBL_MIN)
        {
            tau = 0.0;
            beta = c0;
#pragma unroll
            for (int i = 0; i < 3; ++i)
                if (i > k)
                    qr[i][k] = 0.0;
        }
        else
        {
            beta = sqrt(c0 * c0 + tailsq);
            if (c0 >= 0.0)
                beta = -beta;
            const double inv_d = 1.0 / (c0 - beta);
#pragma unroll
            for (int i = 0; i < 3; ++i)
                if (i > k)
                    qr[i][k] *= inv_d;
            tau = (beta - c0) / beta;
        }
        qr[k][k] = beta;
        hco[k] = tau;
        maxpivot = fmax(maxpivot, fabs(beta));
#pragma unroll
        for (int j = 0; j < 3; ++j)
        {
            if (j > k)
            {
                double s = qr[k][j];
#pragma unroll
                for (int i = 0; i < 3; ++i)
                    if (i > k)
                        s += qr[i][k] * qr[i][j];
                s *= tau;
                qr[k][j] -= s;
#pragma unroll
                for (int i = 0; i < 3; ++i)
                    if (i > k)
                        qr[i][j] -= s * qr[i][k];
            }
        }
    }
    int rank = 0;
    const double premult = maxpivot * (DBL_EPSILON * 3.0);
#pragma unroll
    for (int i = 0; i < 3; ++i)
        if (i < nonzero_pivots && fabs(qr[i][i]) > premult)
            rank++;
    if (rank != 3)
        return 0;

#pragma unroll
    for (int c = 0; c < 3; ++c)
    {
        double rhs[3];
#pragma unroll
        for (int i = 0; i < 3; ++i) rhs[i] = (i == c) ? 1.0 : 0.0;
#pragma unroll
        for (int k = 0; k < 3; ++k)
        {
            double s = rhs[k];
#pragma unroll
            for (int i = 0; i < 3; ++i)
                if (i > k)
                    s += qr[i][k] * rhs[i];
            s *= hco[k];
            rhs[k] -= s;
#pragma unroll
            for (int i = 0; i < 3; ++i)
                if (i > k)
                    rhs[i] -= s * qr[i][k];
        }
        rhs[2] = rhs[2] / qr[2][2];
        rhs[1] = (rhs[1] - qr[1][2] * rhs[2]) / qr[1][1];
        rhs[0] = (rhs[0] - qr[0][1] * rhs[1] - qr[0][2] * rhs[2]) / qr[0][0];
#pragma unroll
        for (int i = 0; i < 3; ++i)
        {
#pragma unroll
            for (int r = 0; r < 3; ++r)
                if (perm[i] == r)
                    inv[r][c] = rhs[i];
        }
    }
    return 1;
}

// Minimum eigenvalue of a symmetric 3x3 given by its lower triangle, closed-form trigonometric roots
// after shifting by trace/3 and scaling by max|a_ij| (Eigen 3.4 direct_selfadjoint_eigenvalues).
__device__ __forceinline__ double fj_min_eig3(double a00, double a10, double a11, double a20, double a21, double a22)
{
    const double shift = (a00 + a11 + a22) / 3.0;
    double m00 = a00 - shift, m11 = a11 - shift, m22 = a22 - shift, m10 = a10, m20 = a20, m21 = a21;
    const double scale =
        fmax(fmax(fabs(m00), fmax(fabs(m11), fabs(m22))), fmax(fabs(m10), fmax(fabs(m20), fabs(m21))));
    if (scale > 0.0)
    {
        const double is = 1.0 / scale;
        m00 *= is;
        m11 *= is;
        m22 *= is;
        m10 *= is;
        m20 *= is;
        m21 *= is;
    }
    const double s_inv3 = 1.0 / 3.0;
    const double s_sqrt3 = 1.7320508075688772;
    const double c0 = m00 * m11 * m22 + 2.0 * m10 * m20 * m21 - m00 * m21 * m21 - m11 * m20 * m20 - m22 * m10 * m10;
    const double c1 = m00 * m11 - m10 * m10 + m00 * m22 - m20 * m20 + m11 * m22 - m21 * m21;
    const double c2 = m00 + m11 + m22;
    const double c2_over_3 = c2 * s_inv3;
    double a_over_3 = (c2 * c2_over_3 - c1) * s_inv3;
    a_over_3 = fmax(a_over_3, 0.0);
    const double half_b = 0.5 * (c0 + c2_over_3 * (2.0 * c2_over_3 * c2_over_3 - c1));
    double q = a_over_3 * a_over_3 * a_over_3 - half_b * half_b;
    q = fmax(q, 0.0);
    const double rho = sqrt(a_over_3);
    const double theta = atan2(sqrt(q), half_b) * s_inv3;
    double sin_theta, cos_theta;
    sincos(theta, &sin_theta, &cos_theta);
    const double r0 = c2_over_3 - rho * (cos_theta + s_sqrt3 * sin_theta);
    const double r1 = c2_over_3 - rho * (cos_theta - s_sqrt3 * sin_theta);
    const double r2 = c2_over_3 + 2.0 * rho * cos_theta;
    return fmin(r0, fmin(r1, r2)) * scale + shift;
}

// ---- SIMDIM = 2 (reference src/VarDefs.h:29-41): the same two Eigen algorithms on 2x2 matrices.
// ColPivHouseholderQR<2x2>::isInvertible / inverse, restated like the 3x3 form above (thresholds scale with n = 2).
__device__ __forceinline__ int fj_qr_inverse2(const double (&a)[2][2], double (&inv)[2][2])
{
    double q00 = a[0][0], q01 = a[0][1], q10 = a[1][0], q11 = a[1][1];
    int p0 = 0, p1 = 1;
    const double n0 = q00 * q00 + q10 * q10, n1 = q01 * q01 + q11 * q11;
    const double maxcol = fmax(sqrt(n0), sqrt(n1));
    const double th = maxcol * DBL_EPSILON / 2.0;
    const double thr_helper = th * th;
    int nonzero_pivots = 2;
    /* k = 0: pivot on the larger column (the first one on a tie) */
    double bigsq = n0;
    if (n1 > n0)
    {
        double t = q00; q00 = q01; q01 = t;
        t = q10; q10 = q11; q11 = t;
        p0 = 1; p1 = 0;
        bigsq = n1;
    }
    if (bigsq < thr_helper * 2.0)
        nonzero_pivots = 0;
    double tau0, beta0;
    {
        const double c0 = q00, tailsq = q10 * q10;
        if (tailsq <= DBL_MIN)
        {
            tau0 = 0.0;
            beta0 = c0;
            q10 = 0.0;
        }
        else
        {
            beta0 = sqrt(c0 * c0 + tailsq);
            if (c0 >= 0.0)
                beta0 = -beta0;
            q10 *= 1.0 / (c0 - beta0);
            tau0 = (beta0 - c0) / beta0;
        }
        q00 = beta0;
        double s = q01 + q10 * q11;
        s *= tau0;
        q01 -= s;
        q11 -= s * q10;
    }
    /* k = 1: one column, one row left */
    if (nonzero_pivots == 2 && q11 * q11 < thr_helper * 1.0)
        nonzero_pivots = 1;
    const double maxpivot = fmax(fabs(q00), fabs(q11));
    const double premult = maxpivot * (DBL_EPSILON * 2.0);
    int rank = 0;
    if (0 < nonzero_pivots && fabs(q00) > premult)
        rank++;
    if (1 < nonzero_pivots && fabs(q11) > premult)
        rank++;
    if (rank != 2)
        return 0;
#pragma unroll
    for (int c = 0; c < 2; ++c)
    {
        double r0 = (c == 0) ? 1.0 : 0.0, r1 = (c == 1) ? 1.0 : 0.0;
        double s = r0 + q10 * r1;
        s *= tau0;
        r0 -= s;
        r1 -= s * q10;
        /* the second reflector has no essential part: tau = 0 */
        r1 = r1 / q11;
        r0 = (r0 - q01 * r1) / q00;
        inv[p0][c] = r0;
        inv[p1][c] = r1;
    }
    return 1;
}

// Minimum eigenvalue of a symmetric 2x2 (lower triangle): Eigen 3.4 direct_selfadjoint_eigenvalues<2x2> after the shift by
// trace / 2 and the scaling by max|a_ij|.
__device__ __forceinline__ double fj_min_eig2(double a00, double a10, double a11)
{
    const double shift = (a00 + a11) / 2.0;
    double m00 = a00 - shift, m11 = a11 - shift, m10 = a10;
    const double scale = fmax(fabs(m00), fmax(fabs(m11), fabs(m10)));
    if (scale > 0.0)
    {
        m00 /= scale;
        m11 /= scale;
        m10 /= scale;
    }
    const double t0 = 0.5 * sqrt((m00 - m11) * (m00 - m11) + 4.0 * m10 * m10);
    const double t1 = 0.5 * (m00 + m11);
    return (t1 - t0) * scale + shift;
}
