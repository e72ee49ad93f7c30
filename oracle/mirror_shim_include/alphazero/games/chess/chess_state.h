// Forwarding header (test infrastructure): lets the reference's own unit-test sources, which include the reference's header paths, compile
// against the B200 host mirror's classes of the same names (alphazero-multi-game_b200/host/alphazero_host.hpp).  oracle/run_mirror_unit_tests.sh
#pragma once
#include "alphazero_host.hpp"
