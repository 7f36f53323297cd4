// jacobi_pd.hpp -- drop-in name.  Code written against the reference
// (#include "jacobi_pd.hpp", jacobi_pd::Jacobi<...>::Diagonalize) compiles unchanged when
// this directory comes first on the include path; the work is done on the GPU by libjpd.so.
#ifndef JPD_B200_JACOBI_PD_HPP_
#define JPD_B200_JACOBI_PD_HPP_
#include "jacobi_pd_cuda.hpp"
#endif
