/* Device-resident module chain: the slice of pixelpipe_process_on_GPU() (src/develop/pixelpipe_gpu.c:191-760)
 * that matters for throughput -- upload the first module's input once, hand each module's output to
 * the next as a device buffer (the "borrow the cached vRAM payload" protocol, :219-224,317-328), and
 * read back only the last output (:456-463).  Host pointers in, host pointers out; the modules run
 * through their process_cl() adapters, each followed -- where the node carries blend parameters -- by the blend of the module's
 * output over its input on the device (develop/pixelpipe_gpu.c:364-454 calls dt_develop_blend_process_cl there).  Cache keys and
 * the CPU fallback ladder stay in the reference's own pixelpipe code and are not reproduced here.
 */
#include "dt_surface.h"
#include <stdlib.h>

typedef int (*b200_process_cl_fn)(struct dt_iop_module_t *, const dt_dev_pixelpipe_t *, const dt_dev_pixelpipe_iop_t *,
                                  cl_mem, cl_mem);

typedef struct b200_pipe_node_t
{
  b200_process_cl_fn process_cl;   /* the module's process_cl() */
  struct dt_iop_module_t *module;
  const dt_dev_pixelpipe_iop_t *piece;
  /* blending, develop/blend.c:657-860: NULL = the module has none.  d_form_mask: the raster / drawn mask of roi_out in device memory
   * (the host rasterises forms; NULL without one); d_mask: receives the final mask when the module publishes it as a raster mask */
  const b200_blend_params_t *blend;
  const float *d_form_mask;
  float *d_mask;
} b200_pipe_node_t;

/* provided by libb200iop.so (device memory for the chain; dt_opencl_alloc_device / copy analogues) */
int b200_dev_alloc(void **ptr, size_t bytes);
void b200_dev_free(void *ptr);
int b200_copy_host_to_device(void *d_dst, const void *h_src, size_t bytes, void *stream);
int b200_copy_device_to_host(void *h_dst, const void *d_src, size_t bytes, void *stream);
int b200_stream_synchronize(void *stream);
int b200_stream_create(void **stream);
void b200_stream_destroy(void *stream);
int b200_event_create(void **event);
void b200_event_destroy(void *event);
int b200_event_record(void *event, void *stream);
int b200_stream_wait_event(void *stream, void *event);

static size_t buffer_bytes(const dt_iop_buffer_dsc_t *dsc, const dt_iop_roi_t *roi)
{
  return dsc->bpp * (size_t)roi->width * (size_t)roi->height; /* pixelpipe_hb.c:985 */
}

/* state kept between calls so steady-state runs allocate nothing: two ping-pong device buffers */
typedef struct b200_pipe_buffers_t
{
  void *buf[2];
  size_t cap[2];
} b200_pipe_buffers_t;

b200_pipe_buffers_t *b200_pipe_buffers_new(void) { return calloc(1, sizeof(b200_pipe_buffers_t)); }
void b200_pipe_buffers_free(b200_pipe_buffers_t *b)
{
  if(!b) return;
  for(int k = 0; k < 2; k++) b200_dev_free(b->buf[k]);
  free(b);
}
static int ensure(b200_pipe_buffers_t *b, int k, size_t bytes)
{
  if(b->cap[k] >= bytes) return 0;
  b200_dev_free(b->buf[k]);
  b->buf[k] = NULL;
  b->cap[k] = 0;
  if(b200_dev_alloc(&b->buf[k], bytes)) return 1;
  b->cap[k] = bytes;
  return 0;
}

/* both buffers large enough for every cacheline of the chain (fusion changes which buffer a module lands in) */
static int size_buffers(const b200_pipe_node_t *nodes, int n_nodes, b200_pipe_buffers_t *bufs)
{
  size_t need = 0;
  for(int k = 0; k < n_nodes; k++)
  {
    const size_t bi = buffer_bytes(&nodes[k].piece->dsc_in, &nodes[k].piece->roi_in);
    const size_t bo = buffer_bytes(&nodes[k].piece->dsc_out, &nodes[k].piece->roi_out);
    if(bi > need) need = bi;
    if(bo > need) need = bo;
  }
  return ensure(bufs, 0, need) || ensure(bufs, 1, need);
}

/* ---- pipe-level fusion of the raw front ------------------------------------------------------------------------
 * rawprepare -> temperature -> highlights are three pointwise modules over the mosaic; one after the other they move
 * 26 bytes per sample through HBM, as one pass over the sensor data 8 (b200_rawfront_process_dev, bit-identical).  The
 * decision belongs to the pipe, where the reference's pixelpipe_process_on_GPU walks its nodes: when the chain starts
 * with rawprepare and the following nodes are temperature and/or highlights in clip mode, they run as one launch and
 * the intermediate cachelines are never produced.  Anything the fused entry point refuses (X-Trans, another highlights
 * mode, mask display) runs module by module as before.  Returns the number of nodes consumed (0 = not fused), < 0 on error. */
#include <string.h>
int b200_rawfront_process_dev(const b200_piece_t *rawprepare, const b200_piece_t *temperature, const b200_piece_t *highlights, const void *d_in,
                              void *d_out, void *stream);
int b200_pipe_fusion_enabled = 1; /* set to 0 to run every module on its own (parity tests compare the two) */
static int fuse_raw_front(const dt_dev_pixelpipe_t *pipe, const b200_pipe_node_t *nodes, int n_nodes, void *d_in, void *d_out)
{
  if(!b200_pipe_fusion_enabled || n_nodes < 2 || !nodes[0].module || strcmp(nodes[0].module->op, "rawprepare")) return 0;
  for(int k = 0; k < n_nodes && k < 3; k++)
    if(nodes[k].blend) return 0; /* a blended module keeps its own launch: the blend reads its input and its output */
  b200_piece_t p[3];
  const b200_piece_t *tp = NULL, *hp = NULL;
  int used = 1;
  b200_piece_from_dt(&p[0], nodes[0].module, pipe, nodes[0].piece);
  if(nodes[used].module && !strcmp(nodes[used].module->op, "temperature"))
  {
    b200_piece_from_dt(&p[1], nodes[used].module, pipe, nodes[used].piece);
    tp = &p[1];
    used++;
  }
  if(used < n_nodes && nodes[used].module && !strcmp(nodes[used].module->op, "highlights"))
  {
    b200_piece_from_dt(&p[2], nodes[used].module, pipe, nodes[used].piece);
    hp = &p[2];
    used++;
  }
  if(used < 2) return 0;
  const int rc = b200_rawfront_process_dev(&p[0], tp, hp, d_in, d_out, pipe->stream);
  if(rc == B200_ERR_UNSUPPORTED) return 0;
  return rc ? -1 : used;
}
/* run nodes [0, n_nodes) with ping-pong buffers; *last = index of the buffer holding the final output */
static int run_chain(const dt_dev_pixelpipe_t *pipe, const b200_pipe_node_t *nodes, int n_nodes, b200_pipe_buffers_t *bufs, int *last)
{
  int k = 0, cur = 0;
  const int fused = fuse_raw_front(pipe, nodes, n_nodes, bufs->buf[0], bufs->buf[1]);
  if(fused < 0) return 1;
  if(fused > 0)
  {
    k = fused;
    cur = 1;
  }
  for(; k < n_nodes; k++, cur ^= 1)
  {
    if(!nodes[k].process_cl(nodes[k].module, pipe, nodes[k].piece, bufs->buf[cur], bufs->buf[cur ^ 1])) return 1;
    if(nodes[k].blend)
    { /* anything the device blend does not take (feathering, ...) fails the chain: the caller runs this pipe the reference's way */
      b200_piece_t p;
      b200_piece_from_dt(&p, nodes[k].module, pipe, nodes[k].piece);
      if(b200_blend_process_dev(&p, nodes[k].blend, bufs->buf[cur], bufs->buf[cur ^ 1], nodes[k].d_form_mask, nodes[k].d_mask, pipe->stream)) return 1;
    }
  }
  *last = cur;
  return 0;
}

/* Returns 0 on success (process() convention).  host_in is the first node's input cacheline,
 * host_out the last node's output cacheline. */
int b200_pixelpipe_process_on_gpu(const dt_dev_pixelpipe_t *pipe, const b200_pipe_node_t *nodes, int n_nodes,
                                  b200_pipe_buffers_t *bufs, const void *host_in, void *host_out)
{
  if(!pipe || !nodes || n_nodes < 1 || !bufs || !host_in || !host_out) return 1;
  if(size_buffers(nodes, n_nodes, bufs)) return 1;
  const size_t in_bytes = buffer_bytes(&nodes[0].piece->dsc_in, &nodes[0].piece->roi_in);
  if(b200_copy_host_to_device(bufs->buf[0], host_in, in_bytes, pipe->stream)) return 1;
  int out_buf = 0;
  if(run_chain(pipe, nodes, n_nodes, bufs, &out_buf)) return 1;
  const dt_dev_pixelpipe_iop_t *last = nodes[n_nodes - 1].piece;
  if(b200_copy_device_to_host(host_out, bufs->buf[out_buf], buffer_bytes(&last->dsc_out, &last->roi_out), pipe->stream))
    return 1;
  return b200_stream_synchronize(pipe->stream);
}


/* ---- several frames in flight --------------------------------------------------------------------
 * A batch export or the darkroom's preview + full pipes keep more than one pipe busy per
 * device; the reference gives each pipe its own OpenCL command queue (opencl.c:1641-1725).  Here: `depth` slots,
 * each with its own stream and ping-pong buffers.  Uploads run ahead on the slot's stream, the module chains
 * of successive frames are ordered by an event (modules share the calling thread's device scratch), and the
 * read-back of frame n overlaps upload and compute of frame n+1 -- PCIe is full duplex. */
#define B200_QUEUE_MAX_DEPTH 8
typedef struct b200_pipe_queue_t
{
  int depth;
  unsigned long submitted;
  void *stream[B200_QUEUE_MAX_DEPTH];
  void *compute_done[B200_QUEUE_MAX_DEPTH];
  b200_pipe_buffers_t bufs[B200_QUEUE_MAX_DEPTH];
} b200_pipe_queue_t;

void b200_pipe_queue_free(b200_pipe_queue_t *q)
{
  if(!q) return;
  for(int k = 0; k < q->depth; k++)
  {
    if(q->stream[k]) b200_stream_synchronize(q->stream[k]);
    b200_event_destroy(q->compute_done[k]);
    b200_stream_destroy(q->stream[k]);
    for(int j = 0; j < 2; j++) b200_dev_free(q->bufs[k].buf[j]);
  }
  free(q);
}
b200_pipe_queue_t *b200_pipe_queue_new(int depth)
{
  if(depth < 1 || depth > B200_QUEUE_MAX_DEPTH) return NULL;
  b200_pipe_queue_t *q = calloc(1, sizeof(*q));
  if(!q) return NULL;
  q->depth = depth;
  for(int k = 0; k < depth; k++)
    if(b200_stream_create(&q->stream[k]) || b200_event_create(&q->compute_done[k]))
    {
      b200_pipe_queue_free(q);
      return NULL;
    }
  return q;
}
/* Enqueue one frame; returns a ticket >= 0, or -1.  host_in must stay valid until the upload has run and
 * host_out until b200_pixelpipe_wait(ticket) returns; both should be pinned for the copies to be asynchronous.
 * Submitting into a slot that is still busy waits for that slot's previous frame first. */
long b200_pixelpipe_submit(b200_pipe_queue_t *q, const dt_dev_pixelpipe_t *pipe, const b200_pipe_node_t *nodes, int n_nodes,
                           const void *host_in, void *host_out)
{
  if(!q || !pipe || !nodes || n_nodes < 1 || !host_in || !host_out) return -1;
  const int slot = (int)(q->submitted % (unsigned long)q->depth);
  void *const st = q->stream[slot];
  b200_pipe_buffers_t *const bufs = &q->bufs[slot];
  if(b200_stream_synchronize(st)) return -1;
  if(size_buffers(nodes, n_nodes, bufs)) return -1;
  if(b200_copy_host_to_device(bufs->buf[0], host_in, buffer_bytes(&nodes[0].piece->dsc_in, &nodes[0].piece->roi_in), st)) return -1;
  if(q->submitted > 0)
  { /* one module chain at a time on the device */
    const int prev = (int)((q->submitted - 1) % (unsigned long)q->depth);
    if(prev != slot && b200_stream_wait_event(st, q->compute_done[prev])) return -1;
  }
  dt_dev_pixelpipe_t p = *pipe;
  p.stream = st;
  int out_buf = 0;
  if(run_chain(&p, nodes, n_nodes, bufs, &out_buf)) return -1;
  if(b200_event_record(q->compute_done[slot], st)) return -1;
  const dt_dev_pixelpipe_iop_t *last = nodes[n_nodes - 1].piece;
  if(b200_copy_device_to_host(host_out, bufs->buf[out_buf], buffer_bytes(&last->dsc_out, &last->roi_out), st)) return -1;
  return (long)(q->submitted++);
}
/* Block until the frame of `ticket` is in its host_out.  Returns 0 on success. */
int b200_pixelpipe_wait(b200_pipe_queue_t *q, long ticket)
{
  if(!q || ticket < 0 || (unsigned long)ticket >= q->submitted) return 1;
  if(q->submitted - (unsigned long)ticket > (unsigned long)q->depth) return 0; /* its slot was reused: already waited for */
  return b200_stream_synchronize(q->stream[(unsigned long)ticket % (unsigned long)q->depth]);
}
