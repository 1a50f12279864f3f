#include "common.cuh"
int build_pyramid(mrtx_ctx*) { mrtx_set_error("pyramid: not built yet"); return MRTX_ERR_STATE; }
int launch_trace(mrtx_ctx*, int, int, int, int, unsigned, unsigned) { mrtx_set_error("trace: not built yet"); return MRTX_ERR_STATE; }
int launch_resolve(mrtx_ctx*) { mrtx_set_error("resolve: not built yet"); return MRTX_ERR_STATE; }
