// TEST INFRASTRUCTURE ONLY (oracle): see sam.h (in-memory stand-in for the htslib calls BAM_handler makes).
#pragma once
#include "sam.h"
