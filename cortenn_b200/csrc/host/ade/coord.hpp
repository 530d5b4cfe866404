// include-path compatibility with the reference (e.g. llo/data.hpp:12-13)
#include "ade/ade.hpp"
