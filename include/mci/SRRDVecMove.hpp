// SRRDVecMove lives in TrialMoveInterface.hpp (header kept for source compatibility with the reference include list)
#include "mci/TrialMoveInterface.hpp"
