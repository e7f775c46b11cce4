// The reference's header name (src/mrslam/msg_factory.h) for sources that are compiled unchanged
// against this repository: add -I include/cgm/ref_names.
#include "../msg_factory.hpp"
