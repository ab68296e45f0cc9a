/* ref_probe.c -- OUR code, compiled INTO oracle/_ref/libchafa_ref.so next to the
 * reference's files.  TEST INFRASTRUCTURE ONLY.  Read-only accessors into the
 * reference's internal state so tests and table generators can see what the
 * reference actually built (symbol table order, cells, palettes). */
#include "config.h"
#include "chafa.h"
#include "internal/chafa-private.h"
#include "internal/chafa-canvas-internal.h"

int ref_probe_version (void) { return 1; }
