/*
 * Variable.hpp -- ad4cl::Variable<group_id>, the C++ operator-overloading front end of
 * the host tape (reference: Variable.hpp:18-100).  Same class, members (`var`, static
 * `gs`, conversion to double) and the twelve + - * / overloads; the reference's copy no
 * longer compiles against its own ad4cl.h (it calls plus_vv, struct variable ... removed
 * by the ad_ rename) -- this one uses the current names and adds the math
 * functions.  Host scalar code only; kernels run through ad4cl_cuda.h.
 */
#ifndef VARIABLE_HPP
#define VARIABLE_HPP
#include "ad4cl.h"

namespace ad4cl {

template <int group_id = 0>
class Variable {
public:
    struct ad_variable var;
    static struct ad_gradient_structure* gs; /* one tape per group_id (Variable.hpp:22,39-40) */

    Variable() {
        ad_init_var(gs, &var, 0.0);
    }

    /* an independent variable with a fresh id */
    explicit Variable(double value) {
        ad_init_var(gs, &var, value);
    }

    Variable(const struct ad_variable& v) {
        var = v;
    }

    operator double() const {
        return var.value;
    }
};

template <int group_id>
struct ad_gradient_structure* Variable<group_id>::gs = create_gradient_structure(DEFAULT_ENTRY_SIZE);

#define AD4CL_VARIABLE_BINARY(OP, NAME)                                                              \
    template <int g>                                                                                 \
    inline const Variable<g> operator OP(const Variable<g>& a, const Variable<g>& b) {              \
        return Variable<g>(ad_##NAME(Variable<g>::gs, a.var, b.var));                                \
    }                                                                                                \
    template <int g>                                                                                 \
    inline const Variable<g> operator OP(const Variable<g>& a, const double& b) {                   \
        return Variable<g>(ad_##NAME##_vd(Variable<g>::gs, a.var, b));                               \
    }                                                                                                \
    template <int g>                                                                                 \
    inline const Variable<g> operator OP(const double& a, const Variable<g>& b) {                   \
        return Variable<g>(ad_##NAME##_dv(Variable<g>::gs, a, b.var));                               \
    }
AD4CL_VARIABLE_BINARY(+, plus)
AD4CL_VARIABLE_BINARY(-, minus)
AD4CL_VARIABLE_BINARY(*, times)
AD4CL_VARIABLE_BINARY(/, divide)
#undef AD4CL_VARIABLE_BINARY

#define AD4CL_VARIABLE_UNARY(NAME)                                                                   \
    template <int g>                                                                                 \
    inline const Variable<g> NAME(const Variable<g>& a) {                                            \
        return Variable<g>(ad_##NAME(Variable<g>::gs, a.var));                                       \
    }
AD4CL_VARIABLE_UNARY(cos) AD4CL_VARIABLE_UNARY(sin) AD4CL_VARIABLE_UNARY(tan)
AD4CL_VARIABLE_UNARY(acos) AD4CL_VARIABLE_UNARY(asin) AD4CL_VARIABLE_UNARY(atan)
AD4CL_VARIABLE_UNARY(cosh) AD4CL_VARIABLE_UNARY(sinh) AD4CL_VARIABLE_UNARY(tanh)
AD4CL_VARIABLE_UNARY(exp) AD4CL_VARIABLE_UNARY(log) AD4CL_VARIABLE_UNARY(log10) AD4CL_VARIABLE_UNARY(sqrt)
#undef AD4CL_VARIABLE_UNARY

template <int g>
inline const Variable<g> pow(const Variable<g>& a, const Variable<g>& b) {
    return Variable<g>(ad_pow(Variable<g>::gs, a.var, b.var));
}
template <int g>
inline const Variable<g> pow(const Variable<g>& a, const double& b) {
    return Variable<g>(ad_pow_vd(Variable<g>::gs, a.var, b));
}
template <int g>
inline const Variable<g> pow(const double& a, const Variable<g>& b) {
    return Variable<g>(ad_pow_dv(Variable<g>::gs, a, b.var));
}

} // namespace ad4cl

#endif /* VARIABLE_HPP */
