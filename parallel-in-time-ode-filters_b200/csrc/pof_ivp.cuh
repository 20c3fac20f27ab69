// Vector fields and Jacobians of the built-in initial value problems (reference pof/ivp.py:7-152), evaluated on the
// device inside the fused linearisation kernels.  ids: POF_IVP_* in include/pof_b200.h.
#pragma once
#include "../../include/pof_b200.h"
#include "pof_small.cuh"

namespace pof {

struct IvpParams {
  double p[8];
};
// vector field f and Jacobian J (row-major dxd) of the built-in problems, reference pof/ivp.py
POF_HD bool ivp_eval(int id, const IvpParams& P, const double* y, double* f, double* J) {
  switch (id) {
    case POF_IVP_LOGISTIC:
      f[0] = y[0] * (1.0 - y[0]);
      J[0] = 1.0 - 2.0 * y[0];
      return true;
    case POF_IVP_LOTKAVOLTERRA: {
      const double a = P.p[0], b = P.p[1], c = P.p[2], dd = P.p[3];
      f[0] = a * y[0] - b * y[0] * y[1];
      f[1] = -c * y[1] + dd * y[0] * y[1];
      J[0] = a - b * y[1]; J[1] = -b * y[0];
      J[2] = dd * y[1];    J[3] = -c + dd * y[0];
      return true;
    }
    case POF_IVP_VANDERPOL: {
      const double mu = P.p[0];
      f[0] = y[1];
      f[1] = mu * ((1.0 - y[0] * y[0]) * y[1] - y[0]);
      J[0] = 0.0; J[1] = 1.0;
      J[2] = mu * (-2.0 * y[0] * y[1] - 1.0); J[3] = mu * (1.0 - y[0] * y[0]);
      return true;
    }
    case POF_IVP_FITZHUGHNAGUMO: {
      const double a = P.p[0], b = P.p[1], tinv = P.p[2], l = P.p[3];
      f[0] = y[0] - (y[0] * y[0] * y[0]) / 3.0 - y[1] + l;
      f[1] = tinv * (y[0] + a - b * y[1]);
      J[0] = 1.0 - y[0] * y[0]; J[1] = -1.0;
      J[2] = tinv;              J[3] = -tinv * b;
      return true;
    }
    case POF_IVP_ROBER: {
      const double k1 = P.p[0], k2 = P.p[1], k3 = P.p[2];
      f[0] = -k1 * y[0] + k3 * y[1] * y[2];
      f[1] = k1 * y[0] - k2 * y[1] * y[1] - k3 * y[1] * y[2];
      f[2] = k2 * y[1] * y[1];
      J[0] = -k1; J[1] = k3 * y[2];                     J[2] = k3 * y[1];
      J[3] = k1;  J[4] = -2.0 * k2 * y[1] - k3 * y[2];  J[5] = -k3 * y[1];
      J[6] = 0.0; J[7] = 2.0 * k2 * y[1];               J[8] = 0.0;
      return true;
    }
    case POF_IVP_RIGIDBODY: {
      const double p0 = P.p[0], p1 = P.p[1], p2 = P.p[2];
      f[0] = p0 * y[1] * y[2]; f[1] = p1 * y[0] * y[2]; f[2] = p2 * y[0] * y[1];
      J[0] = 0.0;       J[1] = p0 * y[2]; J[2] = p0 * y[1];
      J[3] = p1 * y[2]; J[4] = 0.0;       J[5] = p1 * y[0];
      J[6] = p2 * y[1]; J[7] = p2 * y[0]; J[8] = 0.0;
      return true;
    }
    case POF_IVP_SEIR: {
      const double p0 = P.p[0], p1 = P.p[1], p2 = P.p[2], p3 = P.p[3];
      const double inf = p1 * y[0] * y[2] / p3;
      f[0] = -inf; f[1] = inf - p0 * y[1]; f[2] = p0 * y[1] - p2 * y[2]; f[3] = p2 * y[2];
      const double i0 = p1 * y[2] / p3, i2 = p1 * y[0] / p3;
      J[0] = -i0;  J[1] = 0.0;  J[2] = -i2;  J[3] = 0.0;
      J[4] = i0;   J[5] = -p0;  J[6] = i2;   J[7] = 0.0;
      J[8] = 0.0;  J[9] = p0;   J[10] = -p2; J[11] = 0.0;
      J[12] = 0.0; J[13] = 0.0; J[14] = p2;  J[15] = 0.0;
      return true;
    }
    case POF_IVP_THREEBODY: {
      const double mu = P.p[0], mp = 1.0 - P.p[0];
      const double a1 = y[0] + mu, a2 = y[0] - mp, y1 = y[1];
      const double r1s = a1 * a1 + y1 * y1, r2s = a2 * a2 + y1 * y1;
      const double r1 = sqrt(r1s), r2 = sqrt(r2s);
      const double i13 = 1.0 / (r1s * r1), i23 = 1.0 / (r2s * r2);
      const double i15 = i13 / r1s, i25 = i23 / r2s;
      f[0] = y[2];
      f[1] = y[3];
      f[2] = y[0] + 2.0 * y[3] - mp * a1 * i13 - mu * a2 * i23;
      f[3] = y1 - 2.0 * y[2] - mp * y1 * i13 - mu * y1 * i23;
      const double cross = 3.0 * mp * a1 * y1 * i15 + 3.0 * mu * a2 * y1 * i25;
      J[0] = 0.0; J[1] = 0.0; J[2] = 1.0; J[3] = 0.0;
      J[4] = 0.0; J[5] = 0.0; J[6] = 0.0; J[7] = 1.0;
      J[8] = 1.0 - mp * (i13 - 3.0 * a1 * a1 * i15) - mu * (i23 - 3.0 * a2 * a2 * i25);
      J[9] = cross; J[10] = 0.0; J[11] = 2.0;
      J[12] = cross;
      J[13] = 1.0 - mp * (i13 - 3.0 * y1 * y1 * i15) - mu * (i23 - 3.0 * y1 * y1 * i25);
      J[14] = -2.0; J[15] = 0.0;
      return true;
    }
    case POF_IVP_HENONHEILES: {
      const double p = P.p[0];
      f[0] = y[2]; f[1] = y[3];
      f[2] = -y[0] - 2.0 * p * y[0] * y[1];
      f[3] = -y[1] - p * (y[0] * y[0] - y[1] * y[1]);
      J[0] = 0.0; J[1] = 0.0; J[2] = 1.0; J[3] = 0.0;
      J[4] = 0.0; J[5] = 0.0; J[6] = 0.0; J[7] = 1.0;
      J[8] = -1.0 - 2.0 * p * y[1]; J[9] = -2.0 * p * y[0];       J[10] = 0.0; J[11] = 0.0;
      J[12] = -2.0 * p * y[0];      J[13] = -1.0 + 2.0 * p * y[1]; J[14] = 0.0; J[15] = 0.0;
      return true;
    }
    default:
      return false;
  }
}


// Lorenz-96 (POF_IVP_LORENZ96), one (step k, component a) of the linearisation H_k = E1 - J_f E0, c_k = J_f y - f(y) at
// y = E0 m_{k+1}:  f_a = (y_{a+1} - y_{a-2}) y_{a-1} - y_a + F (cyclic, d >= 4).  dense != 0: row a of H (n,d,D) and c
// (n,d); else row a of the compact form [J_f (d x d) | c (d)] per step.  Shared by the kernel and the host simulator.
POF_HD void l96_linearize_row(double forcing, long k, int a, int d, int q, double scale0, double scale1, int dense,
                              const double* means_t1, double* H, double* c, double* Jc) {
  const int Q1 = q + 1, D = d * Q1;
  const int ip1 = (a + 1) % d, im1 = (a + d - 1) % d, im2 = (a + d - 2) % d;
  const double* m = means_t1 + k * D;
  const double ya = scale0 * m[a * Q1], yp1 = scale0 * m[ip1 * Q1], ym1 = scale0 * m[im1 * Q1],
               ym2 = scale0 * m[im2 * Q1];
  const double f = (yp1 - ym2) * ym1 - ya + forcing;
  const double jp1 = ym1, jm2 = -ym1, jm1 = yp1 - ym2, ja = -1.0;
  const double ca = jp1 * yp1 + jm2 * ym2 + jm1 * ym1 + ja * ya - f;
  if (dense) {
    double* Hr = H + (k * d + a) * D;
    for (int j = 0; j < D; ++j) Hr[j] = 0.0;
    Hr[ip1 * Q1] = -jp1 * scale0;
    Hr[im2 * Q1] = -jm2 * scale0;
    Hr[im1 * Q1] = -jm1 * scale0;
    Hr[a * Q1] = -ja * scale0;
    Hr[a * Q1 + 1] += scale1;
    c[k * d + a] = ca;
  } else {
    double* o = Jc + k * ((long)d * d + d);
    for (int b = 0; b < d; ++b) o[a * d + b] = 0.0;
    o[a * d + ip1] = jp1;
    o[a * d + im2] = jm2;
    o[a * d + im1] = jm1;
    o[a * d + a] = ja;
    o[(long)d * d + a] = ca;
  }
}

}  // namespace pof
