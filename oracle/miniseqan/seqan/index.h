// TEST INFRASTRUCTURE: stand-in for <seqan/index.h> (see miniseqan.h)
#include "miniseqan.h"
