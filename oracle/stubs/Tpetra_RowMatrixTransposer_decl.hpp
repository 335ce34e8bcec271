// forwards to the single-header stand-in (oracle/stubs/tpetra_stub.hpp); test infrastructure only
#include "tpetra_stub.hpp"
