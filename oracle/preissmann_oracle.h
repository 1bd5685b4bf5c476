/* preissmann_oracle.h - entry points of the CPU oracle (TEST INFRASTRUCTURE; see preissmann_oracle.c).
 * Same structs as the product ABI (include/preissmann_b200.h) so tests feed both identical inputs;
 * all pointers are HOST pointers. */
#ifndef PREISSMANN_ORACLE_H
#define PREISSMANN_ORACLE_H
#include "../include/preissmann_b200.h"
#ifdef __cplusplus
extern "C" {
#endif
int pr_oracle_run(const pr_config* cfg, const pr_geom* geom, const pr_bc* up, const pr_bc* down,
                  const pr_state* ic, const pr_outputs* out);
/* diagnostic: buf [M][levels-1] receives ||R||_2 of the iteration before the accepted one on the next pr_oracle_run
 * calls (NULL = off) */
void pr_oracle_trace_prev_error(double* buf);
int pr_oracle_newton_step(const pr_config* cfg, const pr_geom* geom, const pr_bc* up, const pr_bc* down,
                          int member, int level, const double* h0, const double* q0, const double* h1,
                          const double* q1, double* stage_record, double* R, double* J, double* delta);
int pr_oracle_gvf(const pr_config* cfg, const pr_geom* geom, const double* q0, int64_t q0_stride,
                  const double* downstream_depth, int64_t hd_stride, double* ic_depth, double* ic_flow,
                  int32_t* status);
int pr_oracle_objective(const pr_config* cfg, const double* up_flow, const double* up_depth, double z0,
                        const double* q_query, const double* h_target, int32_t n_query, double* levels_out,
                        double* rmse_out);
int pr_oracle_section_probe(const pr_geom* geom, int node, int member, double g, double h, double Q, double* out);
double pr_oracle_rating_discharge(const pr_rating* r, double stage);
double pr_oracle_rating_dQdz(const pr_rating* r, double stage);
double pr_oracle_brentq_poly(const double* c, int n, double xa, double xb, int* err);
#ifdef __cplusplus
}
#endif
#endif
