// C-callable wrapper around the UNMODIFIED reference filters (GO1/src/Filter/butterworthLPF.{h,cpp},
// butterworth_filter.{h,cpp}), compiled from /root/reference.  Test infrastructure only: pins oracle/filters.c.
#include <string>
#include <iostream>
#define private public
#include <Filter/butterworthLPF.h>
#undef private
#include <Filter/butterworth_filter.h>

extern "C" {
void* ref_lpf_new(double fsampling, double fcutoff) { butterworthLPF* f = new butterworthLPF(); f->init(fsampling, fcutoff); return f; }
void ref_lpf_free(void* h) { delete static_cast<butterworthLPF*>(h); }
double ref_lpf_filter(void* h, double y) { return static_cast<butterworthLPF*>(h)->filter(y); }
void ref_lpf_coefs(void* h, double* c7) {
  butterworthLPF* f = static_cast<butterworthLPF*>(h);
  c7[0] = f->b0; c7[1] = f->b1; c7[2] = f->b2; c7[3] = f->a1; c7[4] = f->a2; c7[5] = f->a; c7[6] = f->ita;
}
void* ref_force_filter_new() { return new ButterworthFilter(); }
void ref_force_filter_free(void* h) { delete static_cast<ButterworthFilter*>(h); }
double ref_force_filter(void* h, double x) { return static_cast<ButterworthFilter*>(h)->ForceFilter(x); }
}
