// TEST INFRASTRUCTURE: stand-in for <seqan/bam_io.h> (see miniseqan.h)
#include "miniseqan.h"
