// Forwarder: the reference includes <stb/stb_image.h>; the only copy in this
// image is the one vendored under the reference's nanort submodule.
#pragma once
#include <plugin/accel_nanort/nanort/examples/common/stb_image.h>
