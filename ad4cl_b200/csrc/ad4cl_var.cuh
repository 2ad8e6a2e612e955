/*
 * ad4cl_var.cuh -- ad4cl::var, the device-side counterpart of the host's
 * ad4cl::Variable<group_id> (reference: Variable.hpp:18-100): operator overloading over the
 * ad_* functions of ad.cuh, so that an objective can be written
 *        var r = a * x + b - y;   out[id] = r * r;
 * instead of nested ad_minus_vd(gs, ad_plus(gs, ad_times_vd(gs, a, x), b), y) calls.  Every
 * operator forwards to exactly one ad_* function: the tape, the partials (including the
 * AD4CL_REFERENCE_QUIRKS switch) and the results are those of the function-call form, bit for bit.
 * Where the host class finds its gradient structure in a static per-group member, a var carries
 * the handle of its thread's tape.
 */
#pragma once

#include "ad.cuh"

namespace ad4cl {

struct var {
    struct ad_variable v;
    struct ad_gradient_structure* gs;

    AD4CL_DEV var() : v{0.0, 0}, gs(nullptr) {}
    AD4CL_DEV var(struct ad_gradient_structure* g, const struct ad_variable& x) : v(x), gs(g) {}
    AD4CL_DEV operator struct ad_variable() const { return v; }   /* out[id] = r; */
    AD4CL_DEV double value() const { return v.value; }
};

#define AD4CL_VAR_BINARY(OP, NAME)                                                                        \
    AD4CL_DEV var operator OP(const var& a, const var& b) { return var(a.gs, ad_##NAME(a.gs, a.v, b.v)); } \
    AD4CL_DEV var operator OP(const var& a, double b) { return var(a.gs, ad_##NAME##_vd(a.gs, a.v, b)); }  \
    AD4CL_DEV var operator OP(double a, const var& b) { return var(b.gs, ad_##NAME##_dv(b.gs, a, b.v)); }
AD4CL_VAR_BINARY(+, plus)
AD4CL_VAR_BINARY(-, minus)
AD4CL_VAR_BINARY(*, times)
AD4CL_VAR_BINARY(/, divide)
#undef AD4CL_VAR_BINARY

AD4CL_DEV var& operator+=(var& a, const var& b) {
    ad_plus_eq(a.gs, &a.v, b.v);
    return a;
}
AD4CL_DEV var& operator+=(var& a, double b) {
    ad_plus_eq_d(a.gs, &a.v, b);
    return a;
}

#define AD4CL_VAR_UNARY(NAME) \
    AD4CL_DEV var NAME(const var& a) { return var(a.gs, ad_##NAME(a.gs, a.v)); }
AD4CL_VAR_UNARY(cos) AD4CL_VAR_UNARY(sin) AD4CL_VAR_UNARY(tan) AD4CL_VAR_UNARY(acos) AD4CL_VAR_UNARY(asin)
AD4CL_VAR_UNARY(atan) AD4CL_VAR_UNARY(cosh) AD4CL_VAR_UNARY(sinh) AD4CL_VAR_UNARY(tanh) AD4CL_VAR_UNARY(exp)
AD4CL_VAR_UNARY(log) AD4CL_VAR_UNARY(log10) AD4CL_VAR_UNARY(sqrt)
#undef AD4CL_VAR_UNARY

AD4CL_DEV var pow(const var& a, const var& b) { return var(a.gs, ad_pow(a.gs, a.v, b.v)); }
AD4CL_DEV var pow(const var& a, double b) { return var(a.gs, ad_pow_vd(a.gs, a.v, b)); }
AD4CL_DEV var pow(double a, const var& b) { return var(b.gs, ad_pow_dv(b.gs, a, b.v)); }

} // namespace ad4cl
