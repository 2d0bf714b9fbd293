// Forwarder, see stb_image.h.
#pragma once
#include <plugin/accel_nanort/nanort/examples/common/stb_image_write.h>
