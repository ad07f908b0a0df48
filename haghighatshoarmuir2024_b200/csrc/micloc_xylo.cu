// micloc_xylo.cu -- Xylo-quantised integer chain (placeholder until the kernel lands).
#include <cuda_runtime.h>

#include "micloc_common.h"

using namespace micloc;

extern "C" int micloc_xylo_create(const micloc_xylo_config *, int, micloc_xylo **out) {
    if (out) *out = nullptr;
    return set_error(MICLOC_ERR_UNSUPPORTED, "xylo path not built yet");
}
extern "C" int micloc_xylo_destroy(micloc_xylo *) { return MICLOC_OK; }
extern "C" int micloc_xylo_run(micloc_xylo *, const void *, int, int64_t, int64_t, int8_t *, uint8_t *, int32_t *,
                               int32_t *, int32_t *, int32_t, int32_t *, void *) {
    return set_error(MICLOC_ERR_UNSUPPORTED, "xylo path not built yet");
}
extern "C" int micloc_xylo_process(micloc_xylo *, const int8_t *, int64_t, int64_t, uint8_t *, int32_t *, void *) {
    return set_error(MICLOC_ERR_UNSUPPORTED, "xylo path not built yet");
}
