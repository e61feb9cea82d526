// reference-compatible include name -> mcb200 facade (types live in layer.hpp)
#include "../layer.hpp"
