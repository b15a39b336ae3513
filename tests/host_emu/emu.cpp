/*
 * emu.cpp — TEST SCAFFOLDING: compiles the per-trajectory device code of
 * pde_controller_b200/csrc/pdeb_core.h with g++ (a "warp" of one lane) so that its logic can be
 * checked against the oracle on a box without a GPU.  Not a product path: the shipped library
 * has no CPU route, and nothing outside tests/ builds or loads this file.
 */
#include <cstdlib>
#include <cstring>
#include <vector>

#include "../../pde_controller_b200/csrc/pdeb_core.h"

template <int KIND, int N, typename real>
static void run_one(const PdebDevBatch& b, const PdebConventions& cv, const PdebDevOut& out, int64_t s) {
    std::vector<unsigned char> smem(pdeb_clause_smem_bytes<real>(1, b.pmax) + 64);
    std::vector<typename PdebVec2<real>::type> table(N + 1);
    std::vector<real> scales(2 * N + 2), mult(N + 1);
    PdebClauseSlots<real, 1> slots = pdeb_clause_slots<real, 1>(smem.data(), b.pmax, 0);
    double pwl_slot[2] = {0.0, 0.0};
    pdeb_run_reg<KIND, N, real, 1>(b, cv, out, s, s, table.data(), scales.data(), slots, true, pwl_slot, mult.data());
}

extern "C" int emu_eval_batch(const PdebBatch* in, const PdebConventions* conv, int precision, double* rob,
                              double* clause, int32_t* status) {
    PdebDevBatch b;
    b.n = in->n; b.stride = in->stride; b.kmax = in->kmax; b.pmax = in->pmax; b.tmax = in->tmax; b.pad = 0;
    b.kind = in->kind; b.n_elem = in->n_elem; b.L = in->L; b.dt = in->dt; b.T = in->T; b.g0 = in->g0;
    b.x_a = in->x_a; b.x_b = in->x_b; b.rho_in = in->rho_in; b.rho_out = in->rho_out; b.E_in = in->E_in; b.E_out = in->E_out;
    b.n_knots = in->n_knots; b.knot_t = in->knot_t; b.knot_y = in->knot_y; b.n_pred = in->n_pred;
    b.pred_xlo = in->pred_xlo; b.pred_xhi = in->pred_xhi; b.pred_m = in->pred_m; b.pred_b = in->pred_b;
    b.pred_tlo = in->pred_tlo; b.pred_thi = in->pred_thi; b.pred_rel = in->pred_rel; b.pred_top = in->pred_top;
    b.n_tree = in->n_tree; b.tree_op = in->tree_op; b.tree_lhs = in->tree_lhs; b.tree_rhs = in->tree_rhs;
    PdebDevOut out{rob, clause, status};
    PdebConventions cv = *conv;
    for (int64_t s = 0; s < in->n; ++s) {
        int kind = in->kind[s], N = in->n_elem[s];
        if (kind == PDEB_KIND_HEAT && N == 30) {
            if (precision == 32) run_one<PDEB_KIND_HEAT, 30, float>(b, cv, out, s);
            else run_one<PDEB_KIND_HEAT, 30, double>(b, cv, out, s);
        } else if (kind == PDEB_KIND_WAVE && N == 20) {
            if (precision == 32) run_one<PDEB_KIND_WAVE, 20, float>(b, cv, out, s);
            else run_one<PDEB_KIND_WAVE, 20, double>(b, cv, out, s);
        } else {
            rob[s] = PDEB_FAILED_ROBUSTNESS;
            if (status) status[s] = PDEB_ST_BADPARAM;
        }
    }
    return 0;
}
