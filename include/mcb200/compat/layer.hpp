// reference-compatible include name -> mcb200 facade
#include "../layer.hpp"
