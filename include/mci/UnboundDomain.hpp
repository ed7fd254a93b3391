// UnboundDomain lives in DomainInterface.hpp (header kept for source compatibility with the reference include list)
#include "mci/DomainInterface.hpp"
