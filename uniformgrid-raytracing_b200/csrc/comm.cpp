/* comm.cpp -- NCCL plumbing for the image-strip sharding: one communicator per context, the strips
 * are gathered to rank 0 INSIDE the frame (captured in the frame's CUDA graph), straight from the buffer the shade
 * kernel wrote.  NCCL has no gather primitive: rank 0 posts one ncclRecv per peer, every other rank one ncclSend,
 * inside a group.  libnccl is opened lazily with dlopen, so the library loads (and the single-GPU path runs) on
 * machines without NCCL; in a torch process this resolves to the NCCL torch already loaded.
 */
#include "ugrt_internal.h"
#include <dlfcn.h>
#include <string.h>

typedef struct ncclComm *ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
typedef int ncclResult_t;
enum { ncclInt8 = 0, ncclUint8 = 1 };

static struct {
	void *handle;
	ncclResult_t (*GetUniqueId)(ncclUniqueId *);
	ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int);
	ncclResult_t (*CommDestroy)(ncclComm_t);
	ncclResult_t (*Send)(const void *, size_t, int, int, ncclComm_t, cudaStream_t);
	ncclResult_t (*Recv)(void *, size_t, int, int, ncclComm_t, cudaStream_t);
	ncclResult_t (*AllGather)(const void *, void *, size_t, int, ncclComm_t, cudaStream_t);
	ncclResult_t (*GroupStart)(void);
	ncclResult_t (*GroupEnd)(void);
	const char *(*GetErrorString)(ncclResult_t);
} N;

static bool load_nccl(ug_context *ctx)
{
	if (N.handle) return true;
	void *h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
	if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
	if (!h) { ug_fail(ctx, UG_ERR_CUDA, "cannot open libnccl: %s", dlerror()); return false; }
#define SYM(field, name) *(void **)(&N.field) = dlsym(h, name); if (!N.field) { ug_fail(ctx, UG_ERR_CUDA, "libnccl lacks %s", name); return false; }
	SYM(GetUniqueId, "ncclGetUniqueId") SYM(CommInitRank, "ncclCommInitRank") SYM(CommDestroy, "ncclCommDestroy") SYM(Send, "ncclSend")
	SYM(Recv, "ncclRecv") SYM(AllGather, "ncclAllGather") SYM(GroupStart, "ncclGroupStart") SYM(GroupEnd, "ncclGroupEnd")
	SYM(GetErrorString, "ncclGetErrorString")
#undef SYM
	N.handle = h;
	return true;
}

#define UG_NCCL(ctx, call)                                                                                   \
	do {                                                                                                     \
		ncclResult_t r_ = (call);                                                                            \
		if (r_ != 0) return ug_fail(ctx, UG_ERR_CUDA, "%s: %s", #call, N.GetErrorString(r_));                \
	} while (0)

extern "C" int ug_comm_unique_id(void *out128)
{
	if (!out128) return UG_ERR_ARG;
	if (!load_nccl(nullptr)) return UG_ERR_CUDA;
	ncclUniqueId id;
	if (N.GetUniqueId(&id) != 0) return UG_ERR_CUDA;
	memcpy(out128, &id, sizeof(id));
	return UG_OK;
}

extern "C" int ug_comm_init(ug_context *ctx, int rank, int world, const void *unique_id128)
{
	if (!ctx || !unique_id128 || world < 1 || rank < 0 || rank >= world) return ug_fail(ctx, UG_ERR_ARG, "ug_comm_init: bad arguments");
	if (!load_nccl(ctx)) return UG_ERR_CUDA;
	UG_CUDA(ctx, cudaSetDevice(ctx->device));
	if (ctx->comm) { N.CommDestroy((ncclComm_t)ctx->comm); ctx->comm = nullptr; }
	ncclUniqueId id;
	memcpy(&id, unique_id128, sizeof(id));
	ncclComm_t comm = nullptr;
	UG_NCCL(ctx, N.CommInitRank(&comm, world, id, rank));
	ctx->comm = comm;
	ctx->comm_rank = rank;
	ctx->comm_world = world;
	if (ctx->graph_exec) { cudaGraphExecDestroy(ctx->graph_exec); ctx->graph_exec = nullptr; }
	return UG_OK;
}

void ugc_destroy(ug_context *ctx)
{
	if (ctx->comm && N.handle) N.CommDestroy((ncclComm_t)ctx->comm);
	ctx->comm = nullptr;
}

/* strips -> rank 0's full image.  bytes[r] / offs[r] describe every rank's strip inside the full image. */
int ugc_gather_strips(ug_context *ctx, cudaStream_t s, const unsigned char *strip, unsigned char *full, const size_t *bytes, const size_t *offs)
{
	ncclComm_t comm = (ncclComm_t)ctx->comm;
	const int rank = ctx->comm_rank, world = ctx->comm_world;
	if (rank == 0) UG_CUDA(ctx, cudaMemcpyAsync(full + offs[0], strip, bytes[0], cudaMemcpyDeviceToDevice, s));
	UG_NCCL(ctx, N.GroupStart());
	if (rank == 0) {
		for (int r = 1; r < world; r++) UG_NCCL(ctx, N.Recv(full + offs[r], bytes[r], ncclUint8, r, comm, s));
	} else {
		UG_NCCL(ctx, N.Send(strip, bytes[rank], ncclUint8, 0, comm, s));
	}
	UG_NCCL(ctx, N.GroupEnd());
	return UG_OK;
}
