// One compile-time variant of csrc/dgemm.cu as a standalone shared library for tools/dgemm_sweep.py.
#include <stdarg.h>
#include <stdio.h>
#include "../desolver_b200/csrc/dgemm.cu"
namespace dsb {
void set_error(const char *fmt, ...) { va_list ap; va_start(ap, fmt); vfprintf(stderr, fmt, ap); va_end(ap); fputc('\n', stderr); }
}
extern "C" __attribute__((visibility("default"))) int variant_dgemm(const double *A, const double *B, double *C, long long M,
                                                                    long long N, long long K, int sms, void *stream)
{
    return dsb::launch_dgemm(A, B, C, M, N, K, sms, (cudaStream_t)stream);
}
