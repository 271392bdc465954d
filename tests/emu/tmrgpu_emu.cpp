/*
  tests/emu/tmrgpu_emu.cpp -- TEST INFRASTRUCTURE ONLY.
  Compiles the kernel bodies and host orchestration of tmr_b200/csrc/gpu with
  the serial primitive stand-ins of prim_emu.h, exporting the same tmrgpu_*
  C-ABI.  Used by tests/ (marker: not gpu) to check the LOGIC of every kernel
  body against the oracle before GPU time is spent.  Never part of the
  product: tmr_b200 only ever loads tmr_b200/lib/libtmr_b200.so.
*/
#define TMRGPU_EMU 1
#include "tmrgpu_api.inl"

extern "C" {

int tmrgpu_ctx_create(int device, void *stream, tmrgpu_ctx **out) {
  tmrgpu_ctx *c = new tmrgpu_ctx();
  c->c.device = device;
  c->c.stream = stream;
  c->own_stream = false;
  *out = c;
  return 0;
}

int tmrgpu_ctx_destroy(tmrgpu_ctx *ctx) {
  delete ctx;
  return 0;
}

const char *tmrgpu_build_kind(void) { return "host-emulation(test-only)"; }
}
