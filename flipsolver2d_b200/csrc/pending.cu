// Temporary: stages not yet implemented report FS2D_ERR_STATE.
#include "fs2d_internal.h"
#define PENDING(sig) int sig { ctx->lastError = "stage not implemented yet"; return FS2D_ERR_STATE; }
PENDING(gridViscosity(Ctx *ctx, int *))
