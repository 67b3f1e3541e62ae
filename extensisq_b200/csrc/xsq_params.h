// xsq_params.h -- host side of xsq_rk_solve: validation of the ABI arguments
// and construction of the device parameter block (RkDev).  Header-only so that
// the library (xsq_api.cu) and the host emulation of the kernels used by the
// tests (tests/kernel_host/) build the block with the very same code.
// Validation follows the reference constructors:
//   validate_tol            extensisq/common.py:30-54
//   _init_sc_control        extensisq/common.py:166-185
//   validate_first_step /   scipy/integrate/_ivp/common.py:10-23 (third party)
//   validate_max_step
#pragma once
#include <cmath>
#include <cstring>
#include <limits>
#include <string>
#include <vector>

#include "xsq.h"
#include "xsq_user.h"

namespace xsq {

template <class T>
inline MethodInfo info_of() {
    return MethodInfo{T::S, T::ORDER, T::ORDER2, T::FSAL, T::NPOL,
                      {T::SC_KB1, T::SC_KB2, T::SC_A, T::SC_G}, T::STBRAD, T::TANANG};
}

inline bool method_info(int method, MethodInfo* mi) {
    switch (method) {
        case XSQ_TS5: *mi = info_of<tab::Ts5>(); return true;
        case XSQ_BS5: *mi = info_of<tab::BS5>(); return true;
        case XSQ_CK5: *mi = info_of<tab::CK5>(); return true;
        case XSQ_ME4: *mi = info_of<tab::Me4>(); return true;
        case XSQ_PR7: *mi = info_of<tab::Pr7>(); return true;
        case XSQ_PR8: *mi = info_of<tab::Pr8>(); return true;
        case XSQ_PR9: *mi = info_of<tab::Pr9>(); return true;
        case XSQ_CFMR7OSC: *mi = info_of<tab::CFMR7osc>(); return true;
        case XSQ_CKDISC: *mi = info_of<tab::CKdisc>(); return true;
        case XSQ_FI4N: *mi = info_of<tab::Fi4N>(); return true;
        case XSQ_FI5N: *mi = info_of<tab::Fi5N>(); return true;
        case XSQ_MU5NMB: *mi = info_of<tab::Mu5Nmb>(); return true;
        case XSQ_MR6NN: *mi = info_of<tab::MR6NN>(); return true;
        default: return false;
    }
}

struct RhsInfo { const char* name; int id, n_state, n_param; };
static const RhsInfo kBuiltinRhs[] = {
    {"lorenz63", XSQ_RHS_LORENZ63, 3, 3},
    {"vanderpol", XSQ_RHS_VANDERPOL, 2, 1},
    {"arenstorf", XSQ_RHS_ARENSTORF, 4, 1},
    {"nbody32", XSQ_RHS_NBODY32, 192, 33},
};

inline bool rhs_shape(int rhs, int* n_state, int* n_param) {
    for (const RhsInfo& r : kBuiltinRhs)
        if (r.id == rhs) { *n_state = r.n_state; *n_param = r.n_param; return true; }
    return user_rhs_shape(rhs, n_state, n_param);
}

// Build the device parameter block from the ABI struct; validation follows
// the reference constructors.
inline int build_params(const xsq_rk_args_t* a, RkDev* P, MethodInfo* mi,
                        std::vector<double>* atol_full) {
    if (!a || a->struct_size != (int32_t)sizeof(xsq_rk_args_t)) {
        set_detail("xsq_rk_args_t.struct_size mismatch");
        return XSQ_ERR_ARG;
    }
    if (a->method == XSQ_METHOD_SWAG) {
        *mi = MethodInfo{1, 1, 1, 0, 0, {1.0, 0.0, 0.0, 0.9}, 0.0, 0.0};
    } else if (a->method == XSQ_METHOD_USER) {
        if (!user_tableau_info(mi)) {
            set_detail("no user tableau loaded");
            return XSQ_ERR_ARG;
        }
    } else if (!method_info(a->method, mi)) {
        set_detail("unknown method");
        return XSQ_ERR_ARG;
    }
    int ns = 0, np = 0;
    if (!rhs_shape(a->rhs, &ns, &np)) { set_detail("unknown rhs"); return XSQ_ERR_ARG; }
    if (a->n_state != ns || a->n_param != np) {
        set_detail("n_state/n_param do not match the rhs");
        return XSQ_ERR_ARG;
    }
    if (a->method >= XSQ_FI4N && a->method <= XSQ_MR6NN) {
        if (ns % 2) {                               // common.py:1246-1250
            set_detail("This method is for second order problems and `fun` should have "
                       "signature: [v, a] = fun(t, [x, v]).");
            return XSQ_ERR_ARG;
        }
        if (a->n_eval > 0 || a->events != 0 || ns > XSQ_MAX_LANE_STATE * 0 + 192) {
            set_detail("Runge-Kutta-Nystrom methods: final state only (no t_eval, no events)");
            return XSQ_ERR_UNSUPPORTED;
        }
    }
    if (a->n_lanes < 0) { set_detail("n_lanes < 0"); return XSQ_ERR_ARG; }
    if (a->n_lanes > 0 &&
        (!a->y0 || !a->t_final || !a->y_final || !a->n_accepted ||
         !a->n_rejected || !a->nfev || !a->status ||
         (np > 0 && !a->params))) {
        set_detail("required pointer is NULL");
        return XSQ_ERR_ARG;
    }
    // validate_tol, common.py:30-54
    if (!(a->n_atol == 1 || a->n_atol == ns) || !a->atol) {
        set_detail("`atol` has wrong shape.");
        return XSQ_ERR_ARG;
    }
    if (!(a->rtol >= 0)) { set_detail("`rtol` must be positive."); return XSQ_ERR_ARG; }
    atol_full->resize(ns);
    for (int i = 0; i < ns; ++i) {
        double v = a->atol[a->n_atol == 1 ? 0 : i];
        if (!(v >= 0)) { set_detail("`atol` must be positive."); return XSQ_ERR_ARG; }
        (*atol_full)[i] = std::fmax(v, 0x1.0p-511);          // sqrt(tiny)
    }
    std::memset(P, 0, sizeof(*P));
    P->rtol = std::fmin(std::fmax(a->rtol, 0x1.4p-50), 0.1);  // 10*epsneg
    if (ns <= XSQ_MAX_LANE_STATE)
        for (int i = 0; i < ns; ++i) P->atol[i] = (*atol_full)[i];
    // validate_max_step / validate_first_step (scipy _ivp/common.py:10-23)
    if (!(a->max_step > 0)) { set_detail("`max_step` must be positive."); return XSQ_ERR_ARG; }
    if (a->n_forced == 0 && a->first_step > 0 &&
        a->first_step > std::fabs(a->t_bound - a->t0)) {
        set_detail("`first_step` exceeds bounds.");
        return XSQ_ERR_ARG;
    }
    if (a->n_eval < 0 ||
        (a->n_eval > 0 && a->n_lanes > 0 && (!a->t_eval || !a->y_eval))) {
        set_detail("t_eval / y_eval inconsistent");
        return XSQ_ERR_ARG;
    }
    if (a->events != 0) {
        const int ne = user_events_count(a->events);
        if (ne < 0 || ne != a->n_event_fns || !a->ev_terminal || !a->ev_direction ||
            a->ev_capacity < 1 ||
            (a->n_lanes > 0 && (!a->t_events || !a->y_events || !a->ev_count))) {
            set_detail("events arguments inconsistent");
            return XSQ_ERR_ARG;
        }
        if (a->n_forced > 0 || a->rhs == XSQ_RHS_NBODY32 || a->n_state > XSQ_MAX_LANE_STATE) {
            set_detail("events are not available with forced steps or warp-per-system "
                       "right-hand sides");
            return XSQ_ERR_UNSUPPORTED;
        }
        for (int k = 0; k < ne; ++k)
            if (a->ev_terminal[k] < 0) {
                set_detail("The `terminal` attribute of each event must be a boolean or "
                           "positive integer.");           // ivp.py prepare_events
                return XSQ_ERR_ARG;
            }
    }
    if (a->method == XSQ_CKDISC && a->n_forced > 0) {
        set_detail("CKdisc takes no forced step sequence (cash.py:245-388 has its own step rule)");
        return XSQ_ERR_ARG;
    }
    if (a->n_forced < 0 || (a->n_forced > 0 && !a->h_forced)) {
        set_detail("h_forced inconsistent");
        return XSQ_ERR_ARG;
    }
    // _init_sc_control, common.py:166-185
    const double* sc = a->use_sc_params ? a->sc_params : mi->sc;
    const int order_error = mi->order2 < mi->order ? mi->order2 : mi->order;
    P->err_exp = -1.0 / (order_error + 1);
    P->minbeta1 = sc[0] * P->err_exp;
    P->minbeta2 = sc[1] * P->err_exp;
    P->minalpha = -sc[2];
    P->safety = sc[3];
    P->safety_sc = std::pow(sc[3], sc[0] + sc[1]);
    P->log2n = std::log2((double)ns);
    // the controller in the log2 domain (xsq_rk_core.cuh ctl_factor); the same
    // expressions, in the same order, are in oracle/xsq_oracle.c
    P->ctl.a1s = 0.5 * P->err_exp;
    P->ctl.a0s = std::log2(P->safety) - P->ctl.a1s * P->log2n;
    P->ctl.a1c = 0.5 * P->minbeta1;
    P->ctl.a2c = 0.5 * P->minbeta2;
    P->ctl.a0c = std::log2(P->safety_sc) - (P->ctl.a1c + P->ctl.a2c) * P->log2n;
    P->n_lanes = a->n_lanes;
    P->y0 = a->y0;
    P->params = a->params;
    P->t0 = a->t0;
    P->t_bound = a->t_bound;
    // OdeSolver.__init__, base.py:165
    P->direction = (a->t_bound != a->t0) ? (a->t_bound > a->t0 ? 1.0 : -1.0) : 1.0;
    P->first_step = a->first_step;
    P->first_h = a->first_step_lanes;
    if (a->first_step_lanes) {
        if (a->n_forced > 0) {
            set_detail("first_step_lanes does not apply to forced steps");
            return XSQ_ERR_ARG;
        }
        P->first_step = 0.0;                 // the kernels then read init_h[lane]
    }
    P->max_step = a->max_step;
    P->t_eval = a->t_eval;
    P->y_eval = a->y_eval;
    P->n_eval = a->n_eval;
    P->eval_pitch = (a->n_eval + 3) & ~3;     // rows padded to 32 bytes
    P->h_forced = a->h_forced;
    P->n_forced = a->n_forced;
    P->max_steps = a->max_steps > 0 ? a->max_steps
                                    : std::numeric_limits<int>::max();
    int ip = a->interpolant;
    if (ip == XSQ_INTERP_DEFAULT) ip = XSQ_INTERP_LOW;      // bogacki.py:218
    if (ip < XSQ_INTERP_FREE || ip > XSQ_INTERP_BEST) {
        set_detail("interpolant should be one of: 'best', 'low', 'free'");
        return XSQ_ERR_ARG;
    }
    P->interpolant = (a->method == XSQ_METHOD_SWAG) ? a->reserved0 : ip;
    P->t_final = a->t_final;
    P->y_final = a->y_final;
    P->h_next = a->h_next;
    P->n_acc = a->n_accepted;
    P->n_rej = a->n_rejected;
    P->nfev = a->nfev;
    P->status = a->status;
    P->n_eval_done = a->n_eval_done;
    // _init_stiffness_detection, common.py:150-164
    if (a->nfev_stiff_detect < 0) {
        set_detail("`nfev_stiff_detect` must be a non-negative integer.");
        return XSQ_ERR_ARG;
    }
    P->nfev_stiff_detect = (mi->stbrad > 0.0 && mi->tanang > 0.0 && a->n_forced == 0)
                               ? a->nfev_stiff_detect : 0;
    if (P->nfev_stiff_detect > 0 && P->nfev_stiff_detect / mi->s < 1) P->nfev_stiff_detect = mi->s;
    P->stiff_many_steps = P->nfev_stiff_detect > 0 ? P->nfev_stiff_detect / mi->s : 1;
    P->stiff_flags = a->stiff_flags;
    P->n_events = a->events != 0 ? a->n_event_fns : 0;
    P->ev_capacity = a->ev_capacity;
    for (int k = 0; k < P->n_events; ++k) {
        P->ev_terminal[k] = a->ev_terminal[k];
        P->ev_direction[k] = a->ev_direction[k] > 0 ? 1 : (a->ev_direction[k] < 0 ? -1 : 0);
    }
    P->t_events = a->t_events;
    P->y_events = a->y_events;
    P->ev_count = a->ev_count;
    return XSQ_OK;
}

}  // namespace xsq
