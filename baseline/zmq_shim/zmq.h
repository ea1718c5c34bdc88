/*
 * Declaration-only shim for the libzmq 4.x C API.
 *
 * The image ships libzmq.so.5 inside pyzmq but no zmq.h, and there is no network to
 * fetch one. The reference's ZMQ van (src/zmq_van.h) uses ~20 entry points of the
 * public, ABI-stable libzmq interface; they are declared here so the UNMODIFIED
 * reference sources can be compiled and linked against that shared object for the
 * baseline arm of bench.py. This file contains no reference code.
 */
#ifndef PSLITE_B200_ZMQ_SHIM_H_
#define PSLITE_B200_ZMQ_SHIM_H_
#include <stddef.h>
#include <errno.h>
#ifdef __cplusplus
extern "C" {
#endif

/* libzmq-specific errno values (public ABI) */
#define ZMQ_HAUSNUMERO 156384712
#ifndef EFSM
#define EFSM (ZMQ_HAUSNUMERO + 51)
#endif
#ifndef ENOCOMPATPROTO
#define ENOCOMPATPROTO (ZMQ_HAUSNUMERO + 52)
#endif
#ifndef ETERM
#define ETERM (ZMQ_HAUSNUMERO + 53)
#endif
#ifndef EMTHREAD
#define EMTHREAD (ZMQ_HAUSNUMERO + 54)
#endif

/* socket types */
#define ZMQ_PAIR 0
#define ZMQ_PUB 1
#define ZMQ_SUB 2
#define ZMQ_REQ 3
#define ZMQ_REP 4
#define ZMQ_DEALER 5
#define ZMQ_ROUTER 6
/* context options */
#define ZMQ_IO_THREADS 1
#define ZMQ_MAX_SOCKETS 2
/* socket options */
#define ZMQ_IDENTITY 5
#define ZMQ_LINGER 17
#define ZMQ_ROUTER_MANDATORY 33
/* send / recv flags */
#define ZMQ_DONTWAIT 1
#define ZMQ_SNDMORE 2
/* monitor events (only referenced from commented-out code in the reference) */
#define ZMQ_EVENT_ALL 0xFFFF

typedef struct zmq_msg_t {
#if defined(__GNUC__) || defined(__clang__)
  unsigned char _[64] __attribute__((aligned(sizeof(void *))));
#else
  unsigned char _[64];
#endif
} zmq_msg_t;

typedef void(zmq_free_fn)(void *data, void *hint);

int zmq_errno(void);
const char *zmq_strerror(int errnum);

void *zmq_ctx_new(void);
int zmq_ctx_set(void *context, int option, int optval);
int zmq_ctx_destroy(void *context);

void *zmq_socket(void *context, int type);
int zmq_close(void *s);
int zmq_setsockopt(void *s, int option, const void *optval, size_t optvallen);
int zmq_bind(void *s, const char *addr);
int zmq_connect(void *s, const char *addr);
int zmq_socket_monitor(void *s, const char *addr, int events);

int zmq_msg_init(zmq_msg_t *msg);
int zmq_msg_init_data(zmq_msg_t *msg, void *data, size_t size, zmq_free_fn *ffn, void *hint);
int zmq_msg_send(zmq_msg_t *msg, void *s, int flags);
int zmq_msg_recv(zmq_msg_t *msg, void *s, int flags);
int zmq_msg_close(zmq_msg_t *msg);
void *zmq_msg_data(zmq_msg_t *msg);
size_t zmq_msg_size(const zmq_msg_t *msg);
int zmq_msg_more(const zmq_msg_t *msg);

#ifdef __cplusplus
}
#endif
#endif /* PSLITE_B200_ZMQ_SHIM_H_ */
