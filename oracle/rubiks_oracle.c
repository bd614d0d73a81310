/*
 * TEST INFRASTRUCTURE ONLY -- CPU oracle for the RubiksNet shift hot path (see rubiks_oracle_impl.h).
 * Built by oracle/Makefile into oracle/_build/librubiks_oracle.so and loaded through ctypes by
 * oracle/__init__.py.  Never linked or imported by the product package rubiksnet_b200/.
 */
#include <math.h>
#include <stdint.h>

#define REAL float
#define SFX _f32
#include "rubiks_oracle_impl.h"
#undef REAL
#undef SFX

#define REAL double
#define SFX _f64
#include "rubiks_oracle_impl.h"
#undef REAL
#undef SFX

int oracle_abi_version(void) { return 1; }
