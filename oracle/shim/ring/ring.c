/* TEST INFRASTRUCTURE ONLY (oracle): see ring.h in this directory.
 *
 * Plain-C restatement of the 12 ring-buffer entry points dde calls
 * (call sites: /root/reference/src/dopri.c:53,93,103,185,189,232,460,748-759
 * and src/difeq.c:33,51,60,96,100,112,142,154,229,265).
 */
#include "ring.h"

#include <stdlib.h>
#include <string.h>

static size_t slot_of(const ring_buffer *b, const data_t *p) {
  return (size_t) (p - b->data) / b->stride;
}

static data_t *slot_ptr(const ring_buffer *b, size_t slot) {
  return b->data + slot * b->stride;
}

ring_buffer *ring_buffer_create(size_t size, size_t stride,
                                overflow_action on_overflow) {
  ring_buffer *b = (ring_buffer *) calloc(1, sizeof(ring_buffer));
  b->size = size;
  b->stride = stride;
  b->bytes_data = (size + 1) * stride;
  b->on_overflow = on_overflow;
  b->data = (data_t *) calloc(b->bytes_data, 1);
  b->head = b->data;
  b->tail = b->data;
  return b;
}

void ring_buffer_destroy(ring_buffer *buffer) {
  free(buffer->data);
  free(buffer);
}

size_t ring_buffer_size(const ring_buffer *buffer, bool bytes) {
  return bytes ? buffer->size * buffer->stride : buffer->size;
}

size_t ring_buffer_used(const ring_buffer *buffer, bool bytes) {
  const size_t n_slots = buffer->size + 1;
  const size_t h = slot_of(buffer, buffer->head);
  const size_t t = slot_of(buffer, buffer->tail);
  const size_t used = (h + n_slots - t) % n_slots;
  return bytes ? used * buffer->stride : used;
}

bool ring_buffer_is_empty(const ring_buffer *buffer) {
  return buffer->head == buffer->tail;
}

const void *ring_buffer_tail(const ring_buffer *buffer) {
  return buffer->tail;
}

/* k-th oldest committed entry; NULL when k is not below the used count. */
const void *ring_buffer_tail_offset(const ring_buffer *buffer, size_t offset) {
  if (offset >= ring_buffer_used(buffer, false)) {
    return NULL;
  }
  const size_t n_slots = buffer->size + 1;
  return slot_ptr(buffer, (slot_of(buffer, buffer->tail) + offset) % n_slots);
}

/* k-th newest committed entry (0 = the one just before head). */
const void *ring_buffer_head_offset(const ring_buffer *buffer, size_t offset) {
  if (offset >= ring_buffer_used(buffer, false)) {
    return NULL;
  }
  const size_t n_slots = buffer->size + 1;
  const size_t h = slot_of(buffer, buffer->head);
  return slot_ptr(buffer, (h + n_slots - 1 - offset % n_slots) % n_slots);
}

static void grow(ring_buffer *b) {
  const size_t used = ring_buffer_used(b, false);
  size_t new_size = (size_t) ((double) b->size * 1.618) + 1;
  if (new_size < b->size + 8) {
    new_size = b->size + 8;
  }
  data_t *fresh = (data_t *) calloc((new_size + 1) * b->stride, 1);
  /* linearise oldest -> newest, then the (uncommitted) head slot */
  const size_t n_slots = b->size + 1;
  const size_t t = slot_of(b, b->tail);
  for (size_t i = 0; i <= used; ++i) {
    memcpy(fresh + i * b->stride, slot_ptr(b, (t + i) % n_slots), b->stride);
  }
  free(b->data);
  b->data = fresh;
  b->size = new_size;
  b->bytes_data = (new_size + 1) * b->stride;
  b->tail = fresh;
  b->head = fresh + used * b->stride;
}

/* Commit the head slot and return the new head. */
void *ring_buffer_head_advance(ring_buffer *buffer) {
  const bool full = ring_buffer_used(buffer, false) == buffer->size;
  if (full && buffer->on_overflow == OVERFLOW_GROW) {
    grow(buffer);
  }
  const size_t n_slots = buffer->size + 1;
  const bool still_full = ring_buffer_used(buffer, false) == buffer->size;
  buffer->head = slot_ptr(buffer, (slot_of(buffer, buffer->head) + 1) % n_slots);
  if (still_full) {
    /* overwrite policy (or zero capacity): the oldest entry is lost */
    buffer->tail = slot_ptr(buffer, (slot_of(buffer, buffer->tail) + 1) % n_slots);
  }
  return buffer->head;
}

/* For a predicate that holds on a prefix of the entries in tail -> head
 * order: the LAST entry on which it holds.  NULL if the buffer is empty, the
 * starting guess is out of range, or the predicate fails at the tail. */
const void *ring_buffer_search_bisect(const ring_buffer *buffer, size_t i,
                                      ring_predicate pred, void *data) {
  const size_t used = ring_buffer_used(buffer, false);
  if (used == 0 || i >= used) {
    return NULL;
  }
  if (!pred(ring_buffer_tail_offset(buffer, 0), data)) {
    return NULL;
  }
  size_t lo = 0, hi = used; /* pred(lo) holds; pred(hi) fails (hi may be == used) */
  if (i > 0 && pred(ring_buffer_tail_offset(buffer, i), data)) {
    lo = i;
  } else if (i > 0) {
    hi = i;
  }
  while (hi - lo > 1) {
    const size_t mid = lo + (hi - lo) / 2;
    if (pred(ring_buffer_tail_offset(buffer, mid), data)) {
      lo = mid;
    } else {
      hi = mid;
    }
  }
  return ring_buffer_tail_offset(buffer, lo);
}

/* Byte copy including head/tail positions; capacities must agree. */
void ring_buffer_mirror(const ring_buffer *src, ring_buffer *dest) {
  memcpy(dest->data, src->data, src->bytes_data);
  dest->head = dest->data + (src->head - src->data);
  dest->tail = dest->data + (src->tail - src->data);
}

/* Copy the n oldest entries, in order, without consuming them. */
const void *ring_buffer_read(const ring_buffer *buffer, void *dest, size_t n) {
  if (n > ring_buffer_used(buffer, false)) {
    return NULL;
  }
  data_t *out = (data_t *) dest;
  for (size_t i = 0; i < n; ++i) {
    memcpy(out + i * buffer->stride, ring_buffer_tail_offset(buffer, i),
           buffer->stride);
  }
  return buffer->tail;
}
