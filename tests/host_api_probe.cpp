// Test helper: runs every host operator of include/ad4cl.h once and prints
// "op a b value id size eid dx0 id0 dx1 id1" (hex floats), for tests/test_host_api.py.
#include <cstdio>
#include <cstring>
#include "ad4cl.h"
#include "Variable.hpp"

static void show(const char* op, double a, double b, struct ad_gradient_structure* gs, struct ad_variable r) {
    const struct ad_entry& e = gs->gradient_stack[gs->stack_current - 1];
    std::printf("%s %a %a %a %d %d %d %a %d %a %d\n", op, a, b, r.value, r.id, e.size, e.id, e.coeff[0].dx,
            e.coeff[0].id, e.size == 2 ? e.coeff[1].dx : 0.0, e.size == 2 ? e.coeff[1].id : 0);
}

int main() {
    const double pts[3][2] = {{0.3, 1.7}, {0.81, 2.25}, {0.05, 1.01}};
    for (auto& p : pts) {
        const double a = p[0], b = p[1];
        struct ad_gradient_structure* gs = create_gradient_structure(256);
        gs->current_variable_id = 101;
        struct ad_variable va = {a, 7}, vb = {b, 9};
#define RESET gs->current_variable_id = 101; gs->stack_current = 3;   /* as the golden probes were taken */
#define BIN(nm) RESET show(#nm, a, b, gs, ad_##nm(gs, va, vb)); RESET show(#nm "_vd", a, b, gs, ad_##nm##_vd(gs, va, b)); \
                RESET show(#nm "_dv", a, b, gs, ad_##nm##_dv(gs, a, vb));
        BIN(plus) BIN(minus) BIN(times) BIN(divide)
#define UN(nm) RESET show(#nm, a, b, gs, ad_##nm(gs, va));
        UN(cos) UN(sin) UN(tan) UN(acos) UN(asin) UN(atan) UN(cosh) UN(sinh) UN(tanh) UN(exp) UN(log) UN(log10) UN(sqrt)
        RESET show("pow", a, b, gs, ad_pow(gs, va, vb));
        RESET show("pow_vd", a, b, gs, ad_pow_vd(gs, va, b));
        RESET show("pow_dv", a, b, gs, ad_pow_dv(gs, a, vb));
        struct ad_variable t = va;
        RESET ad_plus_eq_v(gs, &t, vb);
        show("plus_eq", a, b, gs, t);
        t = va;
        RESET ad_plus_eq_d(gs, &t, b);
        show("plus_eq_d", a, b, gs, t);
        free(gs->gradient_stack);
        free(gs);
    }
    // Variable.hpp: f = a*b + exp(a)/b - 1.5*a and its gradient through compute_gradient
    typedef ad4cl::Variable<3> V;
    V a(2.0), b(3.0);
    V c = a * b + exp(a) / b - 1.5 * a;
    int n = 0;
    double* g = compute_gradient(*V::gs, n);
    std::printf("VARIABLE %a %a %a\n", double(c), g[a.var.id], g[b.var.id]);
    free(g);
    return 0;
}
