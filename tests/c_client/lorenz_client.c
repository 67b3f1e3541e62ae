/* A plain C client of the C ABI (include/xsq.h): what a non-Python host would
 * write.  Integrates N Lorenz-63 trajectories with Ts5 through
 * xsq_rk_solve_host (host buffers in, host buffers out) and prints per-lane
 * results for the test to compare with the Python host layer.
 *   usage: lorenz_client N t_end  ->  lines "i n_acc n_rej nfev status y0 y1 y2" */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "xsq.h"

int main(int argc, char** argv) {
    const long long N = argc > 1 ? atoll(argv[1]) : 64;
    const double t_end = argc > 2 ? atof(argv[2]) : 2.0;
    int32_t rhs = 0, ns = 0, np = 0;
    if (xsq_rhs_builtin("lorenz63", &rhs, &ns, &np) != XSQ_OK || ns != 3 || np != 3) return 2;
    double* y0 = malloc(sizeof(double) * 3 * N);       /* SoA [n_state][n_lanes] */
    double* prm = malloc(sizeof(double) * 3 * N);
    for (long long i = 0; i < N; ++i) {
        y0[0 * N + i] = 1.0 + 0.01 * (double)i;
        y0[1 * N + i] = 1.0;
        y0[2 * N + i] = 20.0 - 0.02 * (double)i;
        prm[0 * N + i] = 10.0;
        prm[1 * N + i] = 28.0;
        prm[2 * N + i] = 8.0 / 3.0;
    }
    double* t_final = malloc(sizeof(double) * N);
    double* y_final = malloc(sizeof(double) * 3 * N);
    int32_t *acc = malloc(4 * N), *rej = malloc(4 * N), *nfev = malloc(4 * N), *st = malloc(4 * N);
    double atol = 1e-9;
    xsq_rk_args_t a;
    memset(&a, 0, sizeof a);
    a.struct_size = (int32_t)sizeof a;
    a.method = XSQ_TS5;
    a.rhs = rhs;
    a.n_state = 3;
    a.n_param = 3;
    a.n_lanes = N;
    a.y0 = y0;
    a.params = prm;
    a.t0 = 0.0;
    a.t_bound = t_end;
    a.rtol = 1e-6;
    a.atol = &atol;
    a.n_atol = 1;
    a.max_step = 1.0 / 0.0;
    a.max_steps = 1000000;
    a.nfev_stiff_detect = 5000;          /* the reference's default */
    a.t_final = t_final;
    a.y_final = y_final;
    a.n_accepted = acc;
    a.n_rejected = rej;
    a.nfev = nfev;
    a.status = st;
    const int rc = xsq_rk_solve_host(&a, 0);
    if (rc != XSQ_OK) {
        fprintf(stderr, "xsq_rk_solve_host: %s (%s)\n", xsq_strerror(rc), xsq_last_error_detail());
        return 1;
    }
    for (long long i = 0; i < N; ++i)
        printf("%lld %d %d %d %d %a %a %a\n", i, acc[i], rej[i], nfev[i], st[i],
               y_final[0 * N + i], y_final[1 * N + i], y_final[2 * N + i]);
    return 0;
}
