// ORACLE - test infrastructure only. Minimal stand-in for <Windows.h> so the reference's asset pipeline sources
// (Plain/src/Common/FileIO.cpp:2 needs only Sleep) compile on Linux where they lie. Not product code.
#pragma once
#include <unistd.h>
static inline void Sleep(unsigned int milliseconds) { usleep(milliseconds * 1000u); }
