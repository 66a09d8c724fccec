// stubs.cu — entry points declared in include/geopolars_b200.h whose kernels are not written yet.
// They fail loudly (GPL_ERR_UNSUPPORTED); nothing falls back to the CPU.
#include "common.cuh"
using namespace gpl;
#define GPL_STUB(name, ...)                                             \
    extern "C" int name(__VA_ARGS__) {                                  \
        set_error(#name ": not implemented in this build");             \
        return GPL_ERR_UNSUPPORTED;                                     \
    }
GPL_STUB(gpl_array_from_wkb, gpl_ctx *, const uint8_t *, const int32_t *, const uint8_t *, int64_t, gpl_array **)
GPL_STUB(gpl_array_to_wkb, gpl_ctx *, const gpl_array *, int32_t *, uint8_t *, int64_t *)
GPL_STUB(gpl_array_import_arrow, gpl_ctx *, const void *, const void *, gpl_array **)
GPL_STUB(gpl_array_export_arrow, gpl_ctx *, const gpl_array *, void *, void *)
GPL_STUB(gpl_export_f64_arrow, const double *, const uint8_t *, int64_t, void *, void *)
GPL_STUB(gpl_export_bool_arrow, const uint8_t *, const uint8_t *, int64_t, void *, void *)
