/* Plain-C use of the C-ABI (no Python, no torch): parameterise a small batch of 3-DOF spline paths with host buffers.
 *   gcc examples/solve_host.c -Iinclude -Ltoppra_b200 -ltoppra_b200 -Wl,-rpath,$PWD/toppra_b200 -lm -o solve_host
 * Replaces, per path, TOPPRA([JointVelocityConstraint, JointAccelerationConstraint], SplineInterpolator(ss, wp),
 * gridpoints, solver_wrapper="seidel").compute_parameterization(0, 0) of the reference. */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>

#include "toppra_b200.h"

int main(void) {
  enum { B = 8, N_WP = 5, DOF = 3, G = 101 };
  double ss[N_WP], grid[G], wp[B][N_WP][DOF], vlim[DOF][2], alim[DOF][2];
  static double K[B][G][2], sd[B][G], u[B][G - 1];
  int status[B];
  unsigned seed = 12345u;
  for (int i = 0; i < N_WP; ++i) ss[i] = (double)i / (N_WP - 1);
  for (int i = 0; i < G; ++i) grid[i] = (double)i / (G - 1);
  grid[G - 1] = 1.0;
  for (int b = 0; b < B; ++b)
    for (int i = 0; i < N_WP; ++i)
      for (int k = 0; k < DOF; ++k) {
        seed = seed * 1664525u + 1013904223u;
        wp[b][i][k] = ((double)(seed >> 8) / (double)(1u << 24)) * 2.0 - 1.0;
      }
  for (int k = 0; k < DOF; ++k) { vlim[k][0] = -2.0; vlim[k][1] = 2.0; alim[k][0] = -5.0; alim[k][1] = 5.0; }
  int rc = tb_solve_velacc_host(0, ss, &wp[0][0][0], B, N_WP, DOF, grid, G, &vlim[0][0], &alim[0][0], 1, 1, NULL, NULL,
                                &K[0][0][0], &sd[0][0], &u[0][0], status);
  if (rc != 0) { fprintf(stderr, "tb_solve_velacc_host failed (%d): %s\n", rc, tb_last_error()); return 1; }
  for (int b = 0; b < B; ++b) {
    double T = 0.0;
    for (int i = 0; i < G - 1; ++i) T += 2.0 * (grid[i + 1] - grid[i]) / (sd[b][i] + sd[b][i + 1]);
    printf("path %d: status %d, duration %.6f s, max sd %.4f\n", b, status[b], T, sqrt(K[b][G / 2][1]));
    if (status[b] != TB_STATUS_OK || !(T > 0.0) || sd[b][0] != 0.0 || sd[b][G - 1] != 0.0) return 2;
  }
  return 0;
}
