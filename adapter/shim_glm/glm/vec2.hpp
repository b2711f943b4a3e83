// TEST INFRASTRUCTURE: see glm/glm.hpp in this directory (minimal glm stand-in)
#pragma once
#include <glm/glm.hpp>
