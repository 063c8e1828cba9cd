// Host-callable launchers of the element / face kernels (implemented in *.cu).
#pragma once
#include <cuda_runtime.h>

#include <string>

#include "diag_face_host.hpp"
#include "forms.hpp"

namespace jots {

// label of the element kernel the calling thread launched last ("form_apply_patch_kernel<3,4,OP_JAC,CF_KL,4,4,128>"), so
// that measurements name the kernel that actually ran (bench.py: roofline.kernel)
const char*& last_element_kernel();
std::string kernel_label(const char* base, int D, int Q, int op, int cf, const char* tail);

// returns 0 on success, non-zero when the (shape, op, cf) combination is not instantiated
int launch_form_apply(int dim, int p, int op, int cf, const FormArgs& a, const TablesRT& t, cudaStream_t st);
int launch_form_apply_slab(int p, int op, int cf, const FormArgs& a, const TablesRT& t, cudaStream_t st);
int launch_form_apply_patch(int p, int op, int cf, const FormArgs& a, const TablesRT& t, cudaStream_t st);
// true when launch_form_apply_patch would pick a kernel that honours FormArgs::dot (the two-phase patch kernel)
bool patch_kernel_fuses_dot(int p, int op, int cf, const FormArgs& a);
// entries per tile of the CG side stream (FormArgs::side_*) that kernel can carry; 0 = it has none
int patch_kernel_side_capacity(int p, int op, int cf, const FormArgs& a);
// the same two questions for the gather-table kernel: tiles of the launch when it honours FormArgs::dot (else 0), entries per
// tile its side stream can carry
int slab_kernel_dot_tiles(int p, int op, int cf, const FormArgs& a, int* side_capacity);
int patch_items(int p, unsigned* out);   // phase-3 items (pencil | visits << 8 | (el << 5 | ij) << 10 | ... << 20), at most 128
void patch_shape(int p, int& px, int& py);   // patch size for order p (0: no patch kernel)
int launch_form_diag_slab(int p, int op, int cf, const FormArgs& a, const TablesRT& t, cudaStream_t st);
int launch_form_diag(int dim, int p, int op, int cf, const FormArgs& a, const TablesRT& t, cudaStream_t st);
int launch_face(int mode, const FaceArgs& a, cudaStream_t st);
int launch_wall_flux(const WallFluxArgs& a, cudaStream_t st);

// per-shape entry points (one translation unit each so they compile in parallel)
#define JOTS_DECL_SHAPE(NAME) int launch_form_apply_##NAME(int op, int cf, const FormArgs& a, const TablesRT& t, cudaStream_t st);
JOTS_DECL_SHAPE(h1) JOTS_DECL_SHAPE(h2) JOTS_DECL_SHAPE(h3) JOTS_DECL_SHAPE(h4)
JOTS_DECL_SHAPE(q1) JOTS_DECL_SHAPE(q2) JOTS_DECL_SHAPE(q3) JOTS_DECL_SHAPE(q4)
#undef JOTS_DECL_SHAPE

}  // namespace jots
