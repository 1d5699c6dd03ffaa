#pragma once
#include "pt_host.h"
int pt_launch_wavefront(pt_ctx ctx, const pt_render_params *p, const pt::LaunchArgs &args) {
    (void)ctx; (void)p; (void)args;
    return pt_fail(1, "wavefront kernel not built yet");
}
