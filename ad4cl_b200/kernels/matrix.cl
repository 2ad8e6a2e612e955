/*
 * matrix.cl -- config 2 kernels in the ad.cl dialect (this repo's own text).
 *
 * matrix_product: C = A*B on n x n ad_variables, one work-item per element of
 * C over a 2-D NDRange, with the operator sequence of the reference's
 * matrixMult kernels (test/matrix/matrixmul.cl:5-28, 31-61): a fresh leaf
 * accumulator, then for every k one ad_times and one in-place ad_plus_eq --
 * 2 n operators per element.  A has ids 0..n^2-1, B ids n^2..2n^2-1
 * (matrix_mul.cpp:84-90).  Indexing as in the reference: A[k + j*widthA],
 * B[k*widthB + i], C[i + widthA*j].
 *
 * matrix_elementwise: the elementwise set of SURVEY.md 8(d) config 2, one
 * work-item per element: op 0: A*B  1: A+B  2: A/B  3: exp(A)  4: log(1+A).
 */
__kernel void matrix_product(__global struct ad_gradient_structure* gs,
        __global struct ad_entry* gradient_stack,
        __global struct ad_variable* A,
        __global struct ad_variable* B,
        __global struct ad_variable* C,
        int widthA, int widthB) {
    ad_init(gs, gradient_stack);
    const int i = get_global_id(0);
    const int j = get_global_id(1);
    if (i >= widthB || j >= widthA) return; /* NDRange padded to whole work-groups */
    struct ad_variable value;
    ad_init_var_g(gs, &value, 0.0);
    for (int k = 0; k < widthA; k++) {
        ad_plus_eq(gs, &value, ad_times(gs, A[k + j * widthA], B[k * widthB + i]));
    }
    C[i + widthA * j] = value;
}

__kernel void matrix_elementwise(__global struct ad_gradient_structure* gs,
        __global struct ad_entry* gradient_stack,
        __global struct ad_variable* A,
        __global struct ad_variable* B,
        __global struct ad_variable* C,
        int width, int op) {
    ad_init(gs, gradient_stack);
    const int i = get_global_id(0);
    const int j = get_global_id(1);
    if (i >= width || j >= width) return;
    const int e = i + width * j;
    struct ad_variable r;
    if (op == 0) r = ad_times(gs, A[e], B[e]);
    else if (op == 1) r = ad_plus(gs, A[e], B[e]);
    else if (op == 2) r = ad_divide(gs, A[e], B[e]);
    else if (op == 3) r = ad_exp(gs, A[e]);
    else r = ad_log(gs, ad_plus_vd(gs, A[e], 1.0));
    C[e] = r;
}
