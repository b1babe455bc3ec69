// TEST INFRASTRUCTURE: stand-in for <seqan/seq_io.h> (see miniseqan.h)
#include "miniseqan.h"
