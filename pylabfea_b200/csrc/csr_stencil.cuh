// csr_stencil.cuh -- closed-form CSR layout of the structured Q4 mesh (Model.mesh numbering, model.py:932-949):
// row 2n+c (dof c of node n = j (NY+1) + k) holds the dofs of all nodes within +-1 in (j, k), ascending, as (ux, uy)
// pairs.  Kernels compute row starts instead of loading rowptr.
#pragma once
#include <stdint.h>

namespace plf {

__device__ __forceinline__ int64_t csr_row_start(int NX, int NY, int64_t row)
{
    const int NnY = NY + 1;
    const int64_t colw = 4ll * (3 * NY + 1);   // entries per unit of "x neighbours" in one node column
    const int64_t n = row >> 1;
    const int c = (int)(row & 1);
    const int j = (int)(n / NnY), k = (int)(n - (int64_t)j * NnY);
    if (j > NX) return colw * (3ll * (NX + 1) - 2);   // row == Ndof: total number of non-zeros
    const int nxc = 1 + (j > 0) + (j < NX), nyc = 1 + (k > 0) + (k < NY);
    const int64_t pj = (j == 0) ? 0 : 3ll * j - 1;
    const int64_t pk = (k == 0) ? 0 : 3ll * k - 1;
    return colw * pj + 4ll * nxc * pk + (int64_t)c * 2 * nxc * nyc;
}

}  // namespace plf
