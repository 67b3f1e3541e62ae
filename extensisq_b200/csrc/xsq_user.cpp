// xsq_user.cpp -- user right-hand sides and user tableaux.
//
// The reference takes any Python callable as `fun` (extensisq/common.py:187)
// and any subclass of RungeKutta with its own A/B/C/E/P as `method`
// (common.py:88-121, docs/Demo_own_RK.ipynb).  On the device both must be
// compile-time constants of the persistent kernel (a function-pointer call per
// stage would serialise the fp64 pipe), so they are compiled at run time:
// the solver headers are embedded in the library (tools/embed.py), the user's
// CUDA source / tableau image is appended, and NVRTC produces a cubin for
// sm_100a that is loaded through the driver API.
//
// libnvrtc and libcuda are dlopen'ed lazily so that libxsq.so itself loads on
// a machine without a driver (the CPU-side ABI tests).
#include <cuda.h>
#include <dlfcn.h>
#include <nvrtc.h>

#include <cstdio>
#include <cstring>
#include <map>
#include <mutex>
#include <string>
#include <vector>

#include "xsq.h"
#include "xsq_user.h"
#include "xsq_rk_fast.cuh"

namespace xsq {

extern const int kNumEmbedded;
extern const char* const kEmbeddedNames[];
extern const char* const kEmbeddedSources[];

namespace {

std::mutex g_mu;

struct UserRhs {
    std::string src, entry;
    int n_state, n_param;
};
std::vector<UserRhs> g_rhs;               // handle = XSQ_RHS_USER_BASE + index
struct UserEvents {
    std::string src, entry;
    int n;
};
std::vector<UserEvents> g_events;         // handle = index + 1

struct UserTab {
    bool loaded = false;
    int generation = 0;
    unsigned long long hash = 0;      // FNV-1a of the tableau image
    xsq_tableau_t t;
    std::string src;
} g_tab;

struct UserPde {
    std::string src, entry;
    int n_param;
    int n_vector = 0;          // > 0: a general system of this many equations (one slab row)
    bool ready = false;
    CUmodule mod = nullptr;
    CUfunction fn[3] = {nullptr, nullptr, nullptr};
    int dev = -1;
};
std::vector<UserPde> g_pde;               // handle = XSQ_PDE_USER_BASE + index

// NVRTC options of EVERY runtime compilation: the AOT build disables FMA
// contraction (Makefile NVFLAGS -fmad=false: only the explicit fma() calls fuse,
// which keeps device code bit-comparable with the C oracle), so must these.
static const char* const kNvrtcOpts[] = {"--gpu-architecture=sm_100a", "--std=c++17",
                                         "--fmad=false", "-lineinfo", "-default-device"};
static constexpr int kNumNvrtcOpts = sizeof(kNvrtcOpts) / sizeof(kNvrtcOpts[0]);

struct Compiled {
    CUmodule mod = nullptr;
    CUfunction fn = nullptr;
    CUfunction init_fn = nullptr;
    CUfunction probe_fn = nullptr;   // null for SWAG modules
    CUfunction evq_fn = nullptr;     // event queue (modules with event functions)
    int occ = 0;
};
std::map<std::string, Compiled> g_cache;

// ---- lazily bound driver / NVRTC entry points ------------------------------
struct Api {
    void* nvrtc = nullptr;
    void* cuda = nullptr;
    nvrtcResult (*CreateProgram)(nvrtcProgram*, const char*, const char*, int,
                                 const char* const*, const char* const*);
    nvrtcResult (*CompileProgram)(nvrtcProgram, int, const char* const*);
    nvrtcResult (*GetCUBINSize)(nvrtcProgram, size_t*);
    nvrtcResult (*GetCUBIN)(nvrtcProgram, char*);
    nvrtcResult (*GetProgramLogSize)(nvrtcProgram, size_t*);
    nvrtcResult (*GetProgramLog)(nvrtcProgram, char*);
    nvrtcResult (*DestroyProgram)(nvrtcProgram*);
    const char* (*GetErrorString)(nvrtcResult);
    CUresult (*ModuleLoadData)(CUmodule*, const void*);
    CUresult (*ModuleGetFunction)(CUfunction*, CUmodule, const char*);
    CUresult (*LaunchKernel)(CUfunction, unsigned, unsigned, unsigned, unsigned,
                             unsigned, unsigned, unsigned, CUstream, void**,
                             void**);
    CUresult (*OccupancyMaxActiveBlocks)(int*, CUfunction, int, size_t);
    CUresult (*CuGetErrorString)(CUresult, const char**);
    CUresult (*FuncSetAttribute)(CUfunction, CUfunction_attribute, int);
} g_api;

template <class F>
bool bind(void* lib, const char* name, F* out) {
    *out = reinterpret_cast<F>(dlsym(lib, name));
    return *out != nullptr;
}

bool load_nvrtc() {
    if (g_api.nvrtc) return true;
    const char* names[] = {"libnvrtc.so.12", "/usr/local/cuda/lib64/libnvrtc.so.12",
                           "libnvrtc.so"};
    for (const char* n : names) {
        g_api.nvrtc = dlopen(n, RTLD_NOW | RTLD_LOCAL);
        if (g_api.nvrtc) break;
    }
    if (!g_api.nvrtc) { set_detail("cannot dlopen libnvrtc.so.12"); return false; }
    void* l = g_api.nvrtc;
    bool ok = bind(l, "nvrtcCreateProgram", &g_api.CreateProgram) &&
              bind(l, "nvrtcCompileProgram", &g_api.CompileProgram) &&
              bind(l, "nvrtcGetCUBINSize", &g_api.GetCUBINSize) &&
              bind(l, "nvrtcGetCUBIN", &g_api.GetCUBIN) &&
              bind(l, "nvrtcGetProgramLogSize", &g_api.GetProgramLogSize) &&
              bind(l, "nvrtcGetProgramLog", &g_api.GetProgramLog) &&
              bind(l, "nvrtcDestroyProgram", &g_api.DestroyProgram) &&
              bind(l, "nvrtcGetErrorString", &g_api.GetErrorString);
    if (!ok) set_detail("libnvrtc is missing a required symbol");
    return ok;
}

bool load_cuda() {
    if (g_api.cuda) return true;
    g_api.cuda = dlopen("libcuda.so.1", RTLD_NOW | RTLD_LOCAL);
    if (!g_api.cuda) { set_detail("cannot dlopen libcuda.so.1 (no driver)"); return false; }
    void* l = g_api.cuda;
    bool ok = bind(l, "cuModuleLoadData", &g_api.ModuleLoadData) &&
              bind(l, "cuModuleGetFunction", &g_api.ModuleGetFunction) &&
              bind(l, "cuLaunchKernel", &g_api.LaunchKernel) &&
              bind(l, "cuOccupancyMaxActiveBlocksPerMultiprocessor",
                   &g_api.OccupancyMaxActiveBlocks) &&
              bind(l, "cuGetErrorString", &g_api.CuGetErrorString) &&
              bind(l, "cuFuncSetAttribute", &g_api.FuncSetAttribute);
    if (!ok) set_detail("libcuda is missing a required symbol");
    return ok;
}

const char* builtin_tab_name(int method) {
    switch (method) {
        case XSQ_TS5: return "Ts5";
        case XSQ_BS5: return "BS5";
        case XSQ_CK5: return "CK5";
        case XSQ_ME4: return "Me4";
        case XSQ_PR7: return "Pr7";
        case XSQ_PR8: return "Pr8";
        case XSQ_PR9: return "Pr9";
        case XSQ_CFMR7OSC: return "CFMR7osc";
        case XSQ_CKDISC: return "CKdisc";
        case XSQ_FI4N: return "Fi4N";
        case XSQ_FI5N: return "Fi5N";
        case XSQ_MU5NMB: return "Mu5Nmb";
        case XSQ_MR6NN: return "MR6NN";
        default: return nullptr;
    }
}
const char* builtin_rhs_name(int rhs) {
    switch (rhs) {
        case XSQ_RHS_LORENZ63: return "Lorenz63";
        case XSQ_RHS_VANDERPOL: return "VanDerPol";
        case XSQ_RHS_ARENSTORF: return "Arenstorf";
        case XSQ_RHS_NBODY32: return "NBody32";
        default: return nullptr;
    }
}

// constexpr accessor (structure) + __constant__ image accessor (values), the
// same pair tools/gen_header.py emits for the built-ins
void append_table(std::string* images, std::string* s, const char* fn,
                  const char* args, const std::string& idx,
                  const std::vector<double>& v) {
    char buf[64];
    std::string vals;
    for (size_t i = 0; i < v.size(); ++i) {
        std::snprintf(buf, sizeof buf, "%a, ", v[i]);
        vals += buf;
    }
    if (v.empty()) vals = "0.0";
    const std::string n = std::to_string(v.size() ? v.size() : 1);
    *images += std::string("static __constant__ __align__(16) double c_UserTab_") + fn + "[" +
               n + "] = {" + vals + "};\n";
    *s += std::string("    XSQ_HD static constexpr double ") + fn + "(" + args +
          ") {\n        constexpr double T[" + n + "] = {" + vals +
          "};\n        return T[" + idx + "];\n    }\n";
    *s += std::string("    __device__ __forceinline__ static double ") + fn +
          "v(" + args + ") { return c_UserTab_" + fn + "[" + idx + "]; }\n";
}

// the same struct layout tools/gen_header.py emits for the built-ins
std::string tableau_source(const xsq_tableau_t& t) {
    const int s = t.n_stages;
    std::string images = "namespace xsq { namespace tab {\n";
    std::string o = "struct UserTab {\n";
    double cdiff = 1.0;                               // common.py:129-137
    for (int i = 0; i < s; ++i)
        for (int j = 0; j < s; ++j) {
            double d = t.C[i] - t.C[j];
            if (d < 0) d = -d;
            if (d != 0.0 && d < cdiff) cdiff = d;
        }
    if (cdiff < 1e-3) cdiff = 1e-3;
    char buf[512];
    std::snprintf(buf, sizeof buf,
                  "    static constexpr int ID = 100, S = %d, ORDER = %d, "
                  "ORDER2 = %d, FSAL = %d, NPOL = %d, VARIANT = GENERIC;\n"
                  "    static constexpr double H_MIN_A = %a;\n"
                  "    static constexpr double STBRAD = %a, TANANG = %a;\n",
                  s, t.order, t.order_secondary, t.E[s] != 0.0 ? 1 : 0,
                  t.n_poly, 10 * 0x1.0p-53 / cdiff, t.stbrad, t.tanang);
    o += buf;
    std::snprintf(buf, sizeof buf,
                  "    static constexpr double SC_KB1 = %a, SC_KB2 = %a, "
                  "SC_A = %a, SC_G = %a;\n",
                  t.sc_params[0], t.sc_params[1], t.sc_params[2],
                  t.sc_params[3]);
    o += buf;
    std::vector<double> v;
    for (int i = 0; i < s; ++i)
        for (int j = 0; j < s; ++j) v.push_back(t.A[i][j]);
    append_table(&images, &o, "a", "int i, int j", "i * " + std::to_string(s) + " + j", v);
    v.assign(t.B, t.B + s);
    append_table(&images, &o, "b", "int i", "i", v);
    v.assign(t.C, t.C + s);
    append_table(&images, &o, "c", "int i", "i", v);
    v.assign(t.E, t.E + s + 1);
    append_table(&images, &o, "e", "int i", "i", v);
    v.clear();
    for (int i = 0; i <= s; ++i)
        for (int k = 0; k < t.n_poly; ++k) v.push_back(t.P[i][k]);
    append_table(&images, &o, "p", "int i, int k",
                 "i * " + std::to_string(t.n_poly > 0 ? t.n_poly : 1) + " + k", v);
    o += "};\n} }\n";
    return images + o;
}

std::string rhs_wrapper(const UserRhs& r) {
    char buf[1400];
    if (r.n_state > XSQ_MAX_LANE_STATE) {
        // one warp per system (xsq_rhs.cuh, WideSystem): the entry returns one component
        std::snprintf(
            buf, sizeof buf,
            "namespace xsq { namespace rhs {\nstruct UserComponent {\n"
            "    __device__ __forceinline__ static double at(int i, double t, const double* y,\n"
            "        const double* p) { return ::%s(i, t, y, p); }\n};\n"
            "typedef WideSystem<%d, %d, UserComponent> User;\n} }\n",
            r.entry.c_str(), r.n_state, r.n_param);
        return buf;
    }
    std::snprintf(
        buf, sizeof buf,
        "namespace xsq { namespace rhs {\nstruct User {\n"
        "    static constexpr int N = %d, NL = %d, NPAR = %d, NPL = %d;\n"
        "    static constexpr bool WARP = false, PADDED = false;\n"
        "    static constexpr int FLOPS = 0;\n"
        "    __device__ __forceinline__ static int comp(int k, int) { return k; }\n"
        "    __device__ __forceinline__ static void load_params(\n"
        "        const double* __restrict__ params, long long sys, long long n_lanes,\n"
        "        int, double (&p)[NPL]) {\n"
        "#pragma unroll\n"
        "        for (int k = 0; k < NPAR; ++k) p[k] = params[k * n_lanes + sys];\n"
        "    }\n"
        "    __device__ __forceinline__ static void f(double t, const double (&y)[NL],\n"
        "        const double (&p)[NPL], double (&dy)[NL]) { ::%s(t, y, p, dy); }\n"
        "};\n} }\n",
        r.n_state, r.n_state, r.n_param, r.n_param > 0 ? r.n_param : 1,
        r.entry.c_str());
    return buf;
}

int minb_for(int s, int nl) {
    const int kd = (s + 1) * nl;
    return kd <= 21 ? 4 : (kd <= 36 ? 3 : 2);
}

// NVRTC source -> loaded CUmodule
int compile_module(const std::string& src, CUmodule* mod) {
    if (!load_nvrtc()) return XSQ_ERR_NVRTC;
    nvrtcProgram prog;
    nvrtcResult r = g_api.CreateProgram(&prog, src.c_str(), "xsq_user.cu", kNumEmbedded,
                                        kEmbeddedSources, kEmbeddedNames);
    if (r != NVRTC_SUCCESS) { set_detail(g_api.GetErrorString(r)); return XSQ_ERR_NVRTC; }
    r = g_api.CompileProgram(prog, kNumNvrtcOpts, kNvrtcOpts);
    if (r != NVRTC_SUCCESS) {
        size_t n = 0;
        g_api.GetProgramLogSize(prog, &n);
        std::string log(n, '\0');
        if (n) g_api.GetProgramLog(prog, &log[0]);
        set_detail(std::string("NVRTC: ") + g_api.GetErrorString(r) + "\n" + log);
        g_api.DestroyProgram(&prog);
        return XSQ_ERR_NVRTC;
    }
    size_t n = 0;
    g_api.GetCUBINSize(prog, &n);
    std::vector<char> cubin(n);
    g_api.GetCUBIN(prog, cubin.data());
    g_api.DestroyProgram(&prog);
    if (!load_cuda()) return XSQ_ERR_CUDA;
    cudaFree(0);
    CUresult cr = g_api.ModuleLoadData(mod, cubin.data());
    if (cr != CUDA_SUCCESS) {
        const char* es = nullptr;
        g_api.CuGetErrorString(cr, &es);
        set_detail(std::string("driver: ") + (es ? es : "?"));
        return XSQ_ERR_CUDA;
    }
    return XSQ_OK;
}

std::string pde_source(const UserPde& u) {
    std::string s = "#include \"xsq_rkc_kernels.cuh\"\n" + u.src + "\n";
    if (u.n_vector > 0)
        s += "namespace xsq { namespace rkc { namespace pde { struct User {\n"
             "  static constexpr bool VECTOR = true;\n"
             "  static constexpr int N = " + std::to_string(u.n_vector) + ";\n"
             "  __device__ __forceinline__ static double at(int i, double t, const double* y,\n"
             "      const double* p) { return ::" + u.entry + "(i, t, y, p); }\n"
             "  __device__ __forceinline__ static double rhs(double, double, double, double,\n"
             "      double, double, double, double, double, const double*) { return 0.0; }\n"
             "};\n} } }\n";
    else
    s += "namespace xsq { namespace rkc { namespace pde { struct User {\n"
         "  static constexpr bool VECTOR = false;\n"
         "  __device__ __forceinline__ static double rhs(double t, double x, double y, double inv_h2,\n"
         "      double c, double n, double s, double w, double e, const double* p) {\n"
         "    return ::" + u.entry + "(t, x, y, inv_h2, c, n, s, w, e, p); }\n};\n} } }\n";
    s += "using namespace xsq::rkc;\n"
         "extern \"C\" __global__ void __launch_bounds__(256) xsq_pde_eval(Slab S, PeerSync ps,\n"
         "    const double* u, const double* up, const double* dn, double t, double* dy) {\n"
         "  eval_body<pde::User>(S, ps, u, up, dn, t, dy); }\n"
         "extern \"C\" __global__ void __launch_bounds__(256) xsq_pde_stage(Slab S, PeerSync ps,\n"
         "    const double* a, const double* up, const double* dn, const double* b, const double* yn,\n"
         "    const double* fn, double* yj, double t, double mu, double nu, double c3, double hmus,\n"
         "    double ajm1) {\n"
         "  stage_body<pde::User>(S, ps, a, up, dn, b, yn, fn, yj, t, mu, nu, c3, hmus, ajm1); }\n"
         "extern \"C\" __global__ void __launch_bounds__(256) xsq_pde_final(Slab S, PeerSync ps,\n"
         "    const double* y, const double* up, const double* dn, const double* yn, const double* fn,\n"
         "    double* f1, double t, double h, double rtol, double atol, double* partial) {\n"
         "  final_body<pde::User>(S, ps, y, up, dn, yn, fn, f1, t, h, rtol, atol, partial); }\n";
    return s;
}

int compile(const std::string& key, const std::string& src, Compiled* out) {
    if (!load_nvrtc()) return XSQ_ERR_NVRTC;
    nvrtcProgram prog;
    nvrtcResult r = g_api.CreateProgram(&prog, src.c_str(), "xsq_user.cu",
                                        kNumEmbedded, kEmbeddedSources,
                                        kEmbeddedNames);
    if (r != NVRTC_SUCCESS) { set_detail(g_api.GetErrorString(r)); return XSQ_ERR_NVRTC; }
    r = g_api.CompileProgram(prog, kNumNvrtcOpts, kNvrtcOpts);
    if (r != NVRTC_SUCCESS) {
        size_t n = 0;
        g_api.GetProgramLogSize(prog, &n);
        std::string log(n, '\0');
        if (n) g_api.GetProgramLog(prog, &log[0]);
        set_detail(std::string("NVRTC: ") + g_api.GetErrorString(r) + "\n" + log);
        g_api.DestroyProgram(&prog);
        return XSQ_ERR_NVRTC;
    }
    size_t n = 0;
    g_api.GetCUBINSize(prog, &n);
    std::vector<char> cubin(n);
    g_api.GetCUBIN(prog, cubin.data());
    g_api.DestroyProgram(&prog);
    if (!load_cuda()) return XSQ_ERR_CUDA;
    cudaFree(0);                           // make the primary context current
    CUresult cr = g_api.ModuleLoadData(&out->mod, cubin.data());
    if (cr == CUDA_SUCCESS)
        cr = g_api.ModuleGetFunction(&out->fn, out->mod, "xsq_user_kernel");
    if (cr == CUDA_SUCCESS)
        cr = g_api.ModuleGetFunction(&out->init_fn, out->mod, "xsq_user_init");
    if (cr == CUDA_SUCCESS &&
        g_api.ModuleGetFunction(&out->probe_fn, out->mod, "xsq_user_probe") != CUDA_SUCCESS)
        out->probe_fn = nullptr;
    if (cr == CUDA_SUCCESS &&
        g_api.ModuleGetFunction(&out->evq_fn, out->mod, "xsq_user_evq") != CUDA_SUCCESS)
        out->evq_fn = nullptr;
    if (cr == CUDA_SUCCESS)
        cr = g_api.OccupancyMaxActiveBlocks(&out->occ, out->fn, 128, 0);
    if (cr != CUDA_SUCCESS) {
        const char* es = nullptr;
        g_api.CuGetErrorString(cr, &es);
        set_detail(std::string("driver: ") + (es ? es : "?"));
        return XSQ_ERR_CUDA;
    }
    (void)key;
    return XSQ_OK;
}

}  // namespace

// Build the translation unit for (method, rhs); exposed for the CPU-side test
// that NVRTC accepts it (no device needed to compile).
// The fast kernel has no dense output of its own: it serves the adaptive solve
// (no t_eval) of a built-in generic pair and thread-per-system right-hand side
// with event functions when the event queue is large enough for every record the
// solve can produce (RkDev::evq_exact).  Steps on which an event may be terminal
// go through Lane::events_slow, out of line.
static int fast_events_variant(int method, int rhs, int events, RkDev* P) {
    if (events == 0 || P->evq_cap <= 0 || !P->evq_exact) return 0;
    double h_min_a = 0.0;
    switch (method) {
        case XSQ_TS5: h_min_a = tab::Ts5::H_MIN_A; break;
        case XSQ_CK5: h_min_a = tab::CK5::H_MIN_A; break;
        case XSQ_ME4: h_min_a = tab::Me4::H_MIN_A; break;
        case XSQ_PR7: h_min_a = tab::Pr7::H_MIN_A; break;
        case XSQ_PR8: h_min_a = tab::Pr8::H_MIN_A; break;
        case XSQ_PR9: h_min_a = tab::Pr9::H_MIN_A; break;
        default: return 0;
    }
    if (rhs == XSQ_RHS_NBODY32) return 0;
    if (rhs >= XSQ_RHS_USER_BASE) {
        const size_t i = (size_t)(rhs - XSQ_RHS_USER_BASE);
        if (i >= g_rhs.size() || g_rhs[i].n_state > XSQ_MAX_LANE_STATE) return 0;
    }
    if (P->n_forced != 0 || P->n_eval != 0 || P->minalpha != 0.0 || P->max_steps != 0x7fffffff ||
        P->n_lanes >= (1LL << 31))
        return 0;
    if (const char* e = getenv("XSQ_NO_FAST"))
        if (e[0] == '1') return 0;
    fast_prepare_h(*P, h_min_a);
    bool terminal = false;
    for (int k = 0; k < P->n_events; ++k) terminal = terminal || P->ev_terminal[k] != 0;
    return (P->nfev_stiff_detect > 0 ? 2 : 1) + (terminal ? 2 : 0);
}

// variant 0: rk_persistent (everything); 1 / 2: rk_fast without / with the
// stiffness diagnosis -- adaptive stepping of a built-in generic pair with
// event functions none of which is terminal, every root located by the event
// queue kernel (xsq_rk_fast.cuh); 3 / 4: the same with terminal events (steps
// that may end the trajectory are resolved outside the stepping loop)
int user_build_source(int method, int rhs, int events, int variant, std::string* src,
                      std::string* key) {
    std::string tabname, rhsname, body;
    if (events != 0) {
        // scipy's `events=`: the functions are device code too; the core header
        // compiles its event machinery in when XSQ_EVENTS_N is defined
        if (events < 1 || (size_t)events > g_events.size()) {
            set_detail("unknown events handle");
            return XSQ_ERR_ARG;
        }
        const UserEvents& e = g_events[(size_t)events - 1];
        if (variant == 1 || variant == 2) body += "#define XSQ_EVENTS_NO_TERMINAL 1\n";
        body += "#define XSQ_EVENTS_N " + std::to_string(e.n) + "\n"
                "__device__ double " + e.entry + "(int, double, const double*, const double*);\n"
                "namespace xsq { __device__ __forceinline__ double user_event(int k, double t,\n"
                "    const double* y, const double* p) { return ::" + e.entry + "(k, t, y, p); } }\n";
    }
    body += variant != 0 ? "#include \"xsq_rk_fast.cuh\"\n" : "#include \"xsq_rk_core.cuh\"\n";
    if (events != 0) body += g_events[(size_t)events - 1].src + "\n";
    int s = 0, nl = 0;
    const bool swag = method == XSQ_METHOD_SWAG;
    if (swag) {
        body += "#include \"xsq_swag_core.cuh\"\n";
        *key = "SWAG";
        s = 17;
    } else if (method == XSQ_METHOD_USER) {
        if (!g_tab.loaded) { set_detail("no user tableau loaded"); return XSQ_ERR_ARG; }
        body += g_tab.src;
        tabname = "UserTab";
        s = g_tab.t.n_stages;
        *key = "T" + std::to_string(g_tab.hash);      // content, not load count
    } else {
        const char* n = builtin_tab_name(method);
        if (!n) return XSQ_ERR_ARG;
        tabname = n;
        MethodInfo mi;
        (void)mi;
        *key = n;
        s = 17;  // conservative default; refined below for built-ins
        static const int ks[] = {6, 7, 6, 5, 10, 13, 17, 9, 6, 5, 6, 9, 6};
        s = ks[method];
    }
    if (rhs >= XSQ_RHS_USER_BASE) {
        const size_t i = (size_t)(rhs - XSQ_RHS_USER_BASE);
        if (i >= g_rhs.size()) { set_detail("unknown rhs handle"); return XSQ_ERR_ARG; }
        const bool wide = g_rhs[i].n_state > XSQ_MAX_LANE_STATE;
        if (wide) body += "#include \"xsq_rhs.cuh\"\n";
        body += g_rhs[i].src + "\n" + rhs_wrapper(g_rhs[i]);
        rhsname = "User";
        nl = wide ? (g_rhs[i].n_state + 31) / 32 : g_rhs[i].n_state;
        *key += "/U" + std::to_string(i);
    } else {
        const char* n = builtin_rhs_name(rhs);
        if (!n) return XSQ_ERR_ARG;
        body += "#include \"xsq_rhs.cuh\"\n";
        rhsname = n;
        nl = rhs == XSQ_RHS_LORENZ63 ? 3 : (rhs == XSQ_RHS_VANDERPOL ? 2 :
             (rhs == XSQ_RHS_ARENSTORF ? 4 : 6));
        *key += std::string("/") + n;
    }
    int minb = minb_for(s, nl);
    // the event machinery adds live state (previous event values, counts): at 128
    // registers the hot loop spills; measured on the Lorenz / Ts5 Poincare workload:
    // 4 / 3 / 2 CTAs per SM -> 126.6 / 92.8 / 101.2 ms
    if (events != 0 && !swag && variant == 0 && minb > 2) --minb;
    if (const char* e = getenv("XSQ_USER_MINB")) {       // tuning knob: CTAs per SM
        const int v = atoi(e);
        if (v >= 1 && v <= 8) minb = v;
    }
    char buf[512];
    if (swag)
        std::snprintf(buf, sizeof buf,
                      "extern \"C\" __global__ void __launch_bounds__(128, 2)\n"
                      "xsq_user_kernel(const xsq::RkDev P) {\n"
                      "    xsq::swag_persistent_body<xsq::rhs::%s>(P);\n}\n",
                      rhsname.c_str());
    else if (variant != 0)
        std::snprintf(buf, sizeof buf,
                      "extern \"C\" __global__ void __launch_bounds__(128, %d)\n"
                      "xsq_user_kernel(const xsq::RkDev P) {\n"
                      "    xsq::rk_fast_body<xsq::tab::%s, xsq::rhs::%s, 128, %s>(P);\n}\n",
                      minb, tabname.c_str(), rhsname.c_str(),
                      (variant == 2 || variant == 4) ? "true" : "false");
    else
        std::snprintf(buf, sizeof buf,
                      "extern \"C\" __global__ void __launch_bounds__(128, %d)\n"
                      "xsq_user_kernel(const xsq::RkDev P) {\n"
                      "    xsq::rk_persistent_body<xsq::tab::%s, xsq::rhs::%s>(P);\n}\n",
                      minb, tabname.c_str(), rhsname.c_str());
    char buf2[256];
    std::snprintf(buf2, sizeof buf2,
                  "extern \"C\" __global__ void __launch_bounds__(128)\n"
                  "xsq_user_init(const xsq::RkDev P) { xsq::ens_init_body<xsq::rhs::%s>(P); }\n",
                  rhsname.c_str());
    char buf3[320] = "";
    if (!swag)
        std::snprintf(buf3, sizeof buf3,
                      "extern \"C\" __global__ void __launch_bounds__(128)\n"
                      "xsq_user_probe(const xsq::RkDev P, int cost, double stbrad, double tanang) {\n"
                      "    xsq::stiff_queue_body<xsq::rhs::%s>(P, cost, stbrad, tanang);\n}\n",
                      rhsname.c_str());
    char buf4[320] = "";
    int evq_minb = 1;                  // CTAs per SM the queue kernel is compiled for (no cap)
    if (const char* e = getenv("XSQ_EVQ_MINB")) {
        const int v = atoi(e);
        if (v >= 1 && v <= 16) evq_minb = v;
    }
    if (!swag && events != 0)
        std::snprintf(buf4, sizeof buf4,
                      "extern \"C\" __global__ void __launch_bounds__(128, %d)\n"
                      "xsq_user_evq(const xsq::RkDev P) {\n"
                      "    xsq::event_queue_body<xsq::tab::%s, xsq::rhs::%s>(P);\n}\n",
                      evq_minb, tabname.c_str(), rhsname.c_str());
    *src = body + buf + buf2 + buf3 + buf4;
    if (events != 0) *key += "/E" + std::to_string(events);
    *key += "/B" + std::to_string(minb) + "/V" + std::to_string(variant) + "/Q" +
            std::to_string(evq_minb);
    return XSQ_OK;
}

bool user_tableau_info(MethodInfo* mi) {
    std::lock_guard<std::mutex> g(g_mu);
    if (!g_tab.loaded) return false;
    const xsq_tableau_t& t = g_tab.t;
    *mi = MethodInfo{t.n_stages, t.order, t.order_secondary,
                     t.E[t.n_stages] != 0.0 ? 1 : 0, t.n_poly,
                     {t.sc_params[0], t.sc_params[1], t.sc_params[2],
                      t.sc_params[3]}, t.stbrad, t.tanang};
    return true;
}

bool user_rhs_shape(int rhs, int* n_state, int* n_param) {
    std::lock_guard<std::mutex> g(g_mu);
    const size_t i = (size_t)(rhs - XSQ_RHS_USER_BASE);
    if (rhs < XSQ_RHS_USER_BASE || i >= g_rhs.size()) return false;
    *n_state = g_rhs[i].n_state;
    *n_param = g_rhs[i].n_param;
    return true;
}

int user_events_count(int events) {
    std::lock_guard<std::mutex> g(g_mu);
    if (events < 1 || (size_t)events > g_events.size()) return -1;
    return g_events[(size_t)events - 1].n;
}

int user_rk_launch(int method, int rhs, int events, const RkDev& P, int cost, double stbrad,
                   double tanang, cudaStream_t st) {
    std::lock_guard<std::mutex> g(g_mu);
    std::string src, key;
    RkDev Pc = P;
    const int variant = fast_events_variant(method, rhs, events, &Pc);
    int rc = user_build_source(method, rhs, events, variant, &src, &key);
    if (rc != XSQ_OK) return rc;
    {   // a CUmodule belongs to the context it was loaded in: one per device
        int kdev = 0;
        cudaGetDevice(&kdev);
        key += "@" + std::to_string(kdev);
    }
    auto it = g_cache.find(key);
    if (it == g_cache.end()) {
        Compiled c;
        rc = compile(key, src, &c);
        if (rc != XSQ_OK) return rc;
        it = g_cache.emplace(key, c).first;
    }
    const Compiled& c = it->second;
    int dev = 0, n_sm = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev);
    long long grid = (long long)n_sm * (c.occ > 0 ? c.occ : 1);
    int ns_user = 0;
    if (rhs >= XSQ_RHS_USER_BASE) ns_user = g_rhs[rhs - XSQ_RHS_USER_BASE].n_state;
    const bool warp = rhs == XSQ_RHS_NBODY32 || ns_user > XSQ_MAX_LANE_STATE;
    const long long per_block = warp ? 4 : 128;          // systems per 128-thread CTA
    const long long want = (P.n_lanes + per_block - 1) / per_block;
    if (want < grid) grid = want;
    if (grid < 1) grid = 1;
    void* args[] = {&Pc};
    prof_mark(0, st);
    {   // initialisation pass (f0 + h_start), thread per lane
        const long long threads = warp ? P.n_lanes * 32 : P.n_lanes;
        const unsigned igrid = (unsigned)((threads + 127) / 128);
        CUresult ci = g_api.LaunchKernel(c.init_fn, igrid ? igrid : 1, 1, 1, 128, 1, 1, 0,
                                         (CUstream)st, args, nullptr);
        count_launch();
        if (ci != CUDA_SUCCESS) { set_detail("cuLaunchKernel(xsq_user_init) failed"); return XSQ_ERR_CUDA; }
    }
    prof_mark(1, st);
    // dense-output staging buffer (xsq_rk_core.cuh::eval_put): 4 x NL x 128
    int ns = 6;    // built-in rhs with a user tableau: NL <= 6
    if (rhs >= XSQ_RHS_USER_BASE) ns = ns_user > XSQ_MAX_LANE_STATE ? (ns_user + 31) / 32 : ns_user;
    const unsigned smem = P.n_eval > 0 ? (unsigned)(sizeof(double) * 4 * ns * 128) : 0;
    if (smem > 40u * 1024u) {
        // 48 KB is the default cap of static + dynamic shared memory; the kernel
        // also holds ~11 KB statically (stiffness state, math tables)
        CUresult ca = g_api.FuncSetAttribute(
            c.fn, CU_FUNC_ATTRIBUTE_MAX_DYNAMIC_SHARED_SIZE_BYTES, (int)smem);
        if (ca != CUDA_SUCCESS) {
            set_detail("cuFuncSetAttribute(MAX_DYNAMIC_SHARED_SIZE_BYTES) failed");
            return XSQ_ERR_CUDA;
        }
    }
    CUresult cr = g_api.LaunchKernel(c.fn, (unsigned)grid, 1, 1, 128, 1, 1, smem,
                                     (CUstream)st, args, nullptr);
    count_launch();
    if (cr != CUDA_SUCCESS) {
        const char* es = nullptr;
        g_api.CuGetErrorString(cr, &es);
        set_detail(std::string("cuLaunchKernel: ") + (es ? es : "?"));
        return XSQ_ERR_CUDA;
    }
    prof_mark(2, st);
    if (P.stiff_q_cap > 0 && c.probe_fn) {   // the queued stiffness probes
        void* pargs[] = {&Pc, &cost, &stbrad, &tanang};
        CUresult cp = g_api.LaunchKernel(c.probe_fn, (unsigned)(n_sm * 8), 1, 1, 128, 1, 1, 0,
                                         (CUstream)st, pargs, nullptr);
        count_launch();
        if (cp != CUDA_SUCCESS) { set_detail("cuLaunchKernel(xsq_user_probe) failed"); return XSQ_ERR_CUDA; }
    }
    if (P.evq_cap > 0) {                     // the queued event roots
        if (!c.evq_fn) { set_detail("event queue kernel missing from the module"); return XSQ_ERR_CUDA; }
        CUresult ce = g_api.LaunchKernel(c.evq_fn, (unsigned)(n_sm * 16), 1, 1, 128, 1, 1, 0,
                                         (CUstream)st, args, nullptr);
        count_launch();
        if (ce != CUDA_SUCCESS) { set_detail("cuLaunchKernel(xsq_user_evq) failed"); return XSQ_ERR_CUDA; }
    }
    prof_mark(3, st);
    return XSQ_OK;
}

int user_pde_vector_size(int pde) {
    std::lock_guard<std::mutex> g(g_mu);
    const size_t i = (size_t)(pde - XSQ_PDE_USER_BASE);
    if (pde < XSQ_PDE_USER_BASE || i >= g_pde.size()) return 0;
    return g_pde[i].n_vector;
}

int user_pde_kernels(int pde, void* fn[3], int* n_param) {
    std::lock_guard<std::mutex> g(g_mu);
    const size_t i = (size_t)(pde - XSQ_PDE_USER_BASE);
    if (pde < XSQ_PDE_USER_BASE || i >= g_pde.size()) { set_detail("unknown pde handle"); return XSQ_ERR_ARG; }
    UserPde& u = g_pde[i];
    int cur_dev = 0;
    cudaGetDevice(&cur_dev);
    if (u.ready && u.dev != cur_dev) u.ready = false;   // a module lives in one context
    if (!u.ready) {
        u.dev = cur_dev;
        int rc = compile_module(pde_source(u), &u.mod);
        if (rc != XSQ_OK) return rc;
        const char* names[3] = {"xsq_pde_eval", "xsq_pde_stage", "xsq_pde_final"};
        for (int k = 0; k < 3; ++k)
            if (g_api.ModuleGetFunction(&u.fn[k], u.mod, names[k]) != CUDA_SUCCESS) {
                set_detail("user pde module lacks a kernel");
                return XSQ_ERR_CUDA;
            }
        u.ready = true;
    }
    for (int k = 0; k < 3; ++k) fn[k] = (void*)u.fn[k];
    *n_param = u.n_param;
    return XSQ_OK;
}

int user_launch(void* fn, unsigned gx, unsigned gy, unsigned bx, unsigned by, void** args,
                cudaStream_t st) {
    return g_api.LaunchKernel((CUfunction)fn, gx, gy, 1, bx, by, 1, 0, (CUstream)st, args,
                              nullptr) == CUDA_SUCCESS ? 0 : -1;
}

}  // namespace xsq

using namespace xsq;

extern "C" {

int xsq_rhs_register_source(const char* cuda_src, const char* entry,
                            int32_t n_state, int32_t n_param,
                            int32_t* rhs_out) {
    if (!cuda_src || !entry || !rhs_out || n_state < 1 ||
        n_state > XSQ_MAX_WARP_STATE || n_param < 0 ||
        (n_state > XSQ_MAX_LANE_STATE && n_param > 16)) {
        set_detail("xsq_rhs_register_source: bad argument (1 <= n_state <= 1024)");
        return XSQ_ERR_ARG;
    }
    std::lock_guard<std::mutex> g(g_mu);
    g_rhs.push_back(UserRhs{cuda_src, entry, n_state, n_param});
    *rhs_out = XSQ_RHS_USER_BASE + (int)g_rhs.size() - 1;
    return XSQ_OK;
}

int xsq_tableau_load(const xsq_tableau_t* tab) {
    if (!tab || tab->n_stages < 1 || tab->n_stages > XSQ_MAX_STAGES - 1 ||
        tab->n_poly < 0 || tab->n_poly > XSQ_MAX_POLY || tab->order < 1 ||
        tab->order_secondary < 1) {
        set_detail("xsq_tableau_load: bad tableau");
        return XSQ_ERR_ARG;
    }
    std::lock_guard<std::mutex> g(g_mu);
    unsigned long long h = 1469598103934665603ULL;
    const unsigned char* b = reinterpret_cast<const unsigned char*>(tab);
    for (size_t i = 0; i < sizeof(*tab); ++i) h = (h ^ b[i]) * 1099511628211ULL;
    if (g_tab.loaded && h == g_tab.hash && std::memcmp(&g_tab.t, tab, sizeof(*tab)) == 0)
        return XSQ_OK;                 // same image: keep the compiled modules
    g_tab.t = *tab;
    g_tab.src = tableau_source(*tab);
    g_tab.loaded = true;
    g_tab.hash = h;
    ++g_tab.generation;
    return XSQ_OK;
}

int xsq_pde_register_source(const char* cuda_src, const char* entry, int32_t n_param,
                            int32_t* pde_out) {
    if (!cuda_src || !entry || !pde_out || n_param < 0) return XSQ_ERR_ARG;
    std::lock_guard<std::mutex> g(g_mu);
    UserPde u;
    u.src = cuda_src;
    u.entry = entry;
    u.n_param = n_param;
    g_pde.push_back(u);
    *pde_out = XSQ_PDE_USER_BASE + (int)g_pde.size() - 1;
    return XSQ_OK;
}

int xsq_pde_register_vector_source(const char* cuda_src, const char* entry, int32_t n_state,
                                   int32_t n_param, int32_t* pde_out) {
    if (!cuda_src || !entry || !pde_out || n_param < 0 || n_state < 1) return XSQ_ERR_ARG;
    std::lock_guard<std::mutex> g(g_mu);
    UserPde u;
    u.src = cuda_src;
    u.entry = entry;
    u.n_param = n_param;
    u.n_vector = n_state;
    g_pde.push_back(u);
    *pde_out = XSQ_PDE_USER_BASE + (int)g_pde.size() - 1;
    return XSQ_OK;
}

/* Compile-only probe (no device needed): does NVRTC accept the translation
 * unit for (method, rhs)?  Used by the CPU-side tests. */
int xsq_events_register_source(const char* cuda_src, const char* entry, int32_t n_events,
                               int32_t* handle_out) {
    if (!cuda_src || !entry || !handle_out || n_events < 1 || n_events > XSQ_MAX_EVENTS) {
        set_detail("xsq_events_register_source: bad argument (1 <= n_events <= 8)");
        return XSQ_ERR_ARG;
    }
    std::lock_guard<std::mutex> g(g_mu);
    g_events.push_back(UserEvents{cuda_src, entry, n_events});
    *handle_out = (int32_t)g_events.size();
    return XSQ_OK;
}

int xsq_user_compile_check(int32_t method, int32_t rhs) {
    return xsq_events_compile_check(method, rhs, 0);
}

int xsq_events_compile_check(int32_t method, int32_t rhs, int32_t events) {
    std::lock_guard<std::mutex> g(g_mu);
    std::string src, key;
    int variant = 0;
    if (const char* e = getenv("XSQ_CHECK_VARIANT")) variant = atoi(e);   // developer aid
    int rc = user_build_source(method, rhs, events, variant, &src, &key);
    if (rc != XSQ_OK) return rc;
    if (!load_nvrtc()) return XSQ_ERR_NVRTC;
    nvrtcProgram prog;
    if (g_api.CreateProgram(&prog, src.c_str(), "xsq_user.cu", kNumEmbedded,
                            kEmbeddedSources, kEmbeddedNames) != NVRTC_SUCCESS)
        return XSQ_ERR_NVRTC;
    nvrtcResult r = g_api.CompileProgram(prog, kNumNvrtcOpts, kNvrtcOpts);
    if (r != NVRTC_SUCCESS) {
        size_t n = 0;
        g_api.GetProgramLogSize(prog, &n);
        std::string log(n, '\0');
        if (n) g_api.GetProgramLog(prog, &log[0]);
        set_detail(std::string("NVRTC: ") + g_api.GetErrorString(r) + "\n" + log);
    } else if (const char* path = getenv("XSQ_DUMP_CUBIN")) {
        // developer aid: keep the cubin so that `cuobjdump -res-usage` / `-sass`
        // can look at a run-time compiled kernel without a GPU
        size_t n = 0;
        g_api.GetCUBINSize(prog, &n);
        std::vector<char> cubin(n);
        g_api.GetCUBIN(prog, cubin.data());
        if (FILE* fh = std::fopen(path, "wb")) {
            std::fwrite(cubin.data(), 1, n, fh);
            std::fclose(fh);
        }
    }
    g_api.DestroyProgram(&prog);
    return r == NVRTC_SUCCESS ? XSQ_OK : XSQ_ERR_NVRTC;
}

}  // extern "C"
