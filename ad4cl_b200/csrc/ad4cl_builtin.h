#pragma once
/* one row per ahead-of-time compiled kernel: name, its entries [recorder: 0 uniform (_u), 1 lane (_m)]
   [accumulator mode 0..2], signature */
struct ad4cl_builtin_row {
    const char* name;
    const void* entry[2][3];
    const char* signature;
};
/* one row per register-tier instance (builtin_static.cu): valid for exactly `nparams` independents bound as
   params(0, nparams) in signature order and, when frozen_arg >= 0, for that integer argument == frozen_value */
struct ad4cl_static_row {
    const char* name;
    const void* entry;
    const char* signature;
    int nparams;
    int frozen_arg;
    int frozen_value;
    int slots;      /* operand slots one work-item records: the instance is only chosen when the configured tape holds them */
};
