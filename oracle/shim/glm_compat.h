// Force-included (-include) before every reference translation unit.
// The only glm in this image is the 0.9.9.0 pre-release vendored under the
// reference's nanort submodule; it still spells the third template parameter
// of glm::vec/glm::mat `precision`, while include/lm/json.h:37 and
// include/lm/serial.h:61 use the released name `qualifier`.  Alias it.
#pragma once
#ifdef __cplusplus
#include <glm/glm.hpp>
namespace glm { using qualifier = precision; }
#endif
