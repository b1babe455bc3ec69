// TEST INFRASTRUCTURE: stand-in for <seqan/arg_parse.h> (see miniseqan.h)
#include "miniseqan.h"
