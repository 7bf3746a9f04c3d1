// placeholder: filled in by the AMG milestone
#include "b200_internal.h"
#define NI(name) B200_FAIL(name ": not implemented yet")
extern "C" int b200_csr_transpose(b200_handle, b200_csr, b200_csr *) { NI("transpose"); }
extern "C" int b200_csr_multiply(b200_handle, b200_csr, b200_csr, b200_csr *) { NI("multiply"); }
extern "C" int b200_amg_create(b200_amg *) { NI("amg"); }
extern "C" int b200_amg_destroy(b200_handle, b200_amg) { NI("amg"); }
extern "C" int b200_amg_set_int(b200_amg, const char *, int) { NI("amg"); }
extern "C" int b200_amg_set_real(b200_amg, const char *, double) { NI("amg"); }
extern "C" int b200_amg_setup(b200_handle, b200_amg, b200_parcsr) { NI("amg"); }
extern "C" int b200_amg_solve(b200_handle, b200_amg, const double *, double *) { NI("amg"); }
extern "C" int b200_amg_num_levels(b200_amg) { return 0; }
extern "C" b200_csr b200_amg_level_A(b200_amg, int) { return nullptr; }
extern "C" b200_csr b200_amg_level_P(b200_amg, int) { return nullptr; }
extern "C" b200_csr b200_amg_level_S(b200_amg, int) { return nullptr; }
extern "C" const int *b200_amg_level_CF(b200_amg, int) { return nullptr; }
extern "C" const double *b200_amg_level_l1(b200_amg, int) { return nullptr; }
extern "C" int b200_amg_setup_times(b200_amg, double *) { NI("amg"); }
extern "C" int b200_strength(b200_handle, b200_csr, double, double, b200_csr *) { NI("strength"); }
extern "C" int b200_pmis(b200_handle, b200_csr, int, int *) { NI("pmis"); }
extern "C" int b200_extpi_interp(b200_handle, b200_csr, b200_csr, const int *, double, int, b200_csr *) { NI("interp"); }
extern "C" int b200_l1_norms(b200_handle, b200_csr, int, double *) { NI("l1"); }
extern "C" int b200_pcg_solve(b200_handle, b200_parcsr, b200_amg, const double *, double *, double, int, int *, double *, double *) { NI("pcg"); }
