/* world-b200 drop-in for externs/WORLD_v2/src/world/matlabfunctions.h:16-141 (the MATLAB-style
 * helpers libworld.a exports).  Same signatures so third-party callers link.
 * Link-compatibility helpers, HOST code: the batched path uses its own device versions
 * (binary-search interp1, the precomputed randn table, the IIR decimation kernels) and never
 * calls these. */
#ifndef WORLD_MATLABFUNCTIONS_H_
#define WORLD_MATLABFUNCTIONS_H_
#include "world/common.h"
#include "world/macrodefinitions.h"
WORLD_BEGIN_C_DECLS
/* replaces W/src/matlabfunctions.cpp:129-134 */
WORLD_API void fftshift(const double *x, int x_length, double *y);
/* replaces W/src/matlabfunctions.cpp:136-155: index[i] = clamp(#{x <= edges[i]}, 1, x_length-1),
 * computed with a cursor that never moves back (edges ascending) */
WORLD_API void histc(const double *x, int x_length, const double *edges, int edges_length,
                     int *index);
/* replaces W/src/matlabfunctions.cpp:157-182: linear interpolation, linear extrapolation outside
 * the knots */
WORLD_API void interp1(const double *x, const double *y, int x_length, const double *xi,
                       int xi_length, double *yi);
/* replaces W/src/matlabfunctions.cpp:184-210 (r = 2..12) */
WORLD_API void decimate(const double *x, int x_length, int r, double *y);
/* replaces W/src/matlabfunctions.cpp:212-214 */
WORLD_API int matlab_round(double x);
/* replaces W/src/matlabfunctions.cpp:216-218 */
WORLD_API void diff(const double *x, int x_length, double *y);
/* replaces W/src/matlabfunctions.cpp:220-241: knots at x + k shift */
WORLD_API void interp1Q(double x, double shift, const double *y, int x_length, const double *xi,
                        int xi_length, double *yi);
/* replace W/src/matlabfunctions.cpp:247-277: the process-wide xorshift128 stream */
WORLD_API double randn(void);
WORLD_API void randn_reseed(void);
/* replaces W/src/matlabfunctions.cpp:279-313: y receives fft_size samples */
WORLD_API void fast_fftfilt(const double *x, int x_length, const double *h, int h_length,
                            int fft_size, const ForwardRealFFT *forward_real_fft,
                            const InverseRealFFT *inverse_real_fft, double *y);
/* replaces W/src/matlabfunctions.cpp:315-325 */
WORLD_API double matlab_std(const double *x, int x_length);
WORLD_END_C_DECLS
#endif
