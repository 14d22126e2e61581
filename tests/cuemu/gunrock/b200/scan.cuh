// tests/cuemu/gunrock/b200/scan.cuh -- TEST INFRASTRUCTURE: the look-back scan is only used by the HOST
// launchers of advance.cuh, which the emulator does not compile (the tests build degree scans themselves).
#pragma once
