// stub: the emulation build (tests/emu/cuda_emu.h) provides what the kernels need
#pragma once
