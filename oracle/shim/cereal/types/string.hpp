// Stand-in (see ../cereal.hpp): container support is a no-op.
#pragma once
#include "../cereal.hpp"
