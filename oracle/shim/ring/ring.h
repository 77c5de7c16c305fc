/* TEST INFRASTRUCTURE ONLY (oracle): functional restatement of the subset of
 * mrc-ide/ring (>= 1.0.6, /root/reference/DESCRIPTION:20) that dde links to.
 *
 * The real source is an un-vendored third-party dependency (src/ring.c:1 is
 * `#include <ring/ring.c>`); it does no floating point, only slot indexing
 * and search.  The semantics below are restated from dde's call sites and
 * from the reference tests that pin them:
 *
 *   - fixed `stride`-byte slots; capacity `size` entries PLUS ONE spare slot,
 *     so `head` (the next write position) is always writable;
 *   - used entries are [tail, head); head itself is never searchable
 *     (src/dopri.c:446 reads the in-progress step straight from ->head);
 *   - head_advance on a full buffer: OVERFLOW_OVERWRITE drops the oldest
 *     entry, OVERFLOW_GROW enlarges the storage geometrically (and may move
 *     ->data, which src/difeq.c:231-239 detects);
 *   - head_offset(0) is the most recently committed entry
 *     (tests/testthat/test-difeq-delay.R:3-24, Fibonacci with n_history = 2).
 */
#ifndef DDE_ORACLE_SHIM_RING_H
#define DDE_ORACLE_SHIM_RING_H

#include <stdbool.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef unsigned char data_t;

typedef enum overflow_action {
  OVERFLOW_OVERWRITE,
  OVERFLOW_GROW,
  OVERFLOW_ERROR
} overflow_action;

typedef struct ring_buffer {
  size_t size;          /* capacity in entries (excluding the spare slot) */
  size_t stride;        /* bytes per entry */
  size_t bytes_data;    /* (size + 1) * stride */
  overflow_action on_overflow;
  data_t *data;         /* base of storage */
  data_t *head;         /* next slot to be written */
  data_t *tail;         /* oldest committed slot */
} ring_buffer;

typedef bool ring_predicate(const void *x, void *data);

ring_buffer *ring_buffer_create(size_t size, size_t stride,
                                overflow_action on_overflow);
void ring_buffer_destroy(ring_buffer *buffer);
size_t ring_buffer_size(const ring_buffer *buffer, bool bytes);
size_t ring_buffer_used(const ring_buffer *buffer, bool bytes);
bool ring_buffer_is_empty(const ring_buffer *buffer);
const void *ring_buffer_tail(const ring_buffer *buffer);
const void *ring_buffer_tail_offset(const ring_buffer *buffer, size_t offset);
const void *ring_buffer_head_offset(const ring_buffer *buffer, size_t offset);
void *ring_buffer_head_advance(ring_buffer *buffer);
const void *ring_buffer_search_bisect(const ring_buffer *buffer, size_t i,
                                      ring_predicate pred, void *data);
void ring_buffer_mirror(const ring_buffer *src, ring_buffer *dest);
const void *ring_buffer_read(const ring_buffer *buffer, void *dest, size_t n);

#ifdef __cplusplus
}
#endif

#endif
