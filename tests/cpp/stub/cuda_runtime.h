// test-harness stub (tests/cpp/h2_host.cc builds the device h2 code for the host): nothing from the CUDA runtime is needed there
#pragma once
