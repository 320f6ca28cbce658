/* comm.cpp -- host side of the one exchange step of the multi-GPU path: strips (or whole frames) are gathered
 * to rank 0 INSIDE the frame (captured in the frame's CUDA graph).  NCCL only bootstraps; the data plane is the library's own
 * kernels over NVLink peer memory.
 *
 * Transports, chosen collectively (allreduce-min of "it worked here") when the destination buffer is (re)allocated:
 *   push   rank 0 exports its gather buffer with a CUDA IPC handle (broadcast over NCCL once) and every other rank maps it.
 *          Default: the shading kernels store every pixel a second time straight into rank 0's frame (P2P stores,
 *          k_shade.cu) and the "ready" / "done" handshakes are flag words in the page behind the image slots
 *          (k_comm_ready / k_comm_done) - "ready" = rank 0's stream is past its read-out of the frame about to be
 *          overwritten, "done" = every rank's stores are complete, and it carries the collective "did any pair arena
 *          overflow" bit.  Variants kept for comparison and as fallbacks: one copy-engine node per rank instead of the fused
 *          stores (UG_NO_FUSED_GATHER=1), 4-byte ncclAllReduce handshakes instead of the flags (UG_NCCL_HANDSHAKE=1).
 *   nccl   when IPC mapping is refused (or UG_NO_PUSH=1): rank 0 posts one ncclRecv per peer, every other rank one
 *          ncclSend, inside a group (NCCL has no gather primitive).
 * The light grids of a strip gather are shared the same way (ugc_share_light: every rank exports one buffer and maps all
 * the others; k_lightshare.cu stores into them).
 * libnccl is opened lazily with dlopen, so the library loads (and the single-GPU path runs) on machines without
 * NCCL; in a torch process this resolves to the NCCL torch already loaded.
 */
#include "ugrt_internal.h"
#include <dlfcn.h>
#include <stdlib.h>
#include <string.h>
#include <vector>

typedef struct ncclComm *ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
typedef int ncclResult_t;
enum { ncclInt8 = 0, ncclUint8 = 1, ncclInt32 = 2, ncclUint32 = 3 };
enum { ncclSum = 0, ncclProd = 1, ncclMax = 2, ncclMin = 3 };

static struct {
	void *handle;
	ncclResult_t (*GetUniqueId)(ncclUniqueId *);
	ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int);
	ncclResult_t (*CommDestroy)(ncclComm_t);
	ncclResult_t (*Send)(const void *, size_t, int, int, ncclComm_t, cudaStream_t);
	ncclResult_t (*Recv)(void *, size_t, int, int, ncclComm_t, cudaStream_t);
	ncclResult_t (*AllGather)(const void *, void *, size_t, int, ncclComm_t, cudaStream_t);
	ncclResult_t (*AllReduce)(const void *, void *, size_t, int, int, ncclComm_t, cudaStream_t);
	ncclResult_t (*Broadcast)(const void *, void *, size_t, int, int, ncclComm_t, cudaStream_t);
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
	SYM(GetErrorString, "ncclGetErrorString") SYM(AllReduce, "ncclAllReduce") SYM(Broadcast, "ncclBroadcast")
#undef SYM
	N.handle = h;
	return true;
}

#define UG_NCCL(ctx, call)                                                                                   \
	do {                                                                                                     \
		ncclResult_t r_ = (call);                                                                            \
		if (r_ != 0) return ug_fail(ctx, UG_ERR_CUDA, "%s: %s", #call, N.GetErrorString(r_));                \
	} while (0)

static void unmap_peer(ug_context *ctx)
{
	if (ctx->peer_full) cudaIpcCloseMemHandle(ctx->peer_full);
	ctx->peer_full = nullptr;
}

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
	if (world > UG_MAX_RANKS) return ug_fail(ctx, UG_ERR_ARG, "ug_comm_init: %d ranks, at most %d are supported", world, UG_MAX_RANKS);
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
	unmap_peer(ctx);
	ctx->use_push = false;
	ctx->flag_handshake = getenv("UG_NCCL_HANDSHAKE") == nullptr;
	ugc_unshare_light(ctx);
	ctx->rgb_full_bytes = 0; /* the gather buffer is shared again (collectively) by the next frame */
	ug_drop_graphs(ctx);
	return UG_OK;
}

void ugc_unshare_light(ug_context *ctx);
void ugc_destroy(ug_context *ctx)
{
	ugc_unshare_light(ctx);
	if (ctx->d_lg_handles) cudaFree(ctx->d_lg_handles);
	ctx->d_lg_handles = nullptr;
	unmap_peer(ctx);
	if (ctx->d_comm_scratch) cudaFree(ctx->d_comm_scratch);
	ctx->d_comm_scratch = nullptr;
	if (ctx->comm && N.handle) N.CommDestroy((ncclComm_t)ctx->comm);
	ctx->comm = nullptr;
}

/* Collective: (re)allocate rank 0's gather buffer and decide the transport.  Every rank of the communicator calls this
 * from the same frame call (the size depends on screen and world size only). */
int ugc_share_full(ug_context *ctx, size_t bytes)
{
	ncclComm_t comm = (ncclComm_t)ctx->comm;
	cudaStream_t s = ctx->stream;
	const int rank = ctx->comm_rank;
	if (!ctx->d_comm_scratch) {
		UG_CUDA(ctx, cudaMalloc((void **)&ctx->d_comm_scratch, 256));
		UG_CUDA(ctx, cudaMemset(ctx->d_comm_scratch, 0, 256));
	}
	unsigned char *scr = (unsigned char *)ctx->d_comm_scratch; /* [0,64) IPC handle, [64,68) ok flags, [128,132) barrier word */
	/* 1. nobody keeps the old buffer mapped when rank 0 frees it */
	UG_CUDA(ctx, cudaStreamSynchronize(s));
	unmap_peer(ctx);
	ctx->use_push = false;
	UG_NCCL(ctx, N.AllReduce(scr + 128, scr + 128, 1, ncclUint32, ncclSum, comm, s));
	UG_CUDA(ctx, cudaStreamSynchronize(s));
	/* 2. rank 0 allocates and exports */
	cudaIpcMemHandle_t h;
	memset(&h, 0, sizeof(h));
	unsigned int ok = 1;
	if (rank == 0) {
		if (ctx->d_rgb_full) cudaFree(ctx->d_rgb_full);
		ctx->d_rgb_full = nullptr;
		UG_CUDA(ctx, cudaMalloc((void **)&ctx->d_rgb_full, bytes + UG_COMM_FLAGS_BYTES));
		UG_CUDA(ctx, cudaMemsetAsync(ctx->d_rgb_full + bytes, 0, UG_COMM_FLAGS_BYTES, s)); /* handshake flags, all sequence numbers 0 */
		if (getenv("UG_NO_PUSH") || cudaIpcGetMemHandle(&h, ctx->d_rgb_full) != cudaSuccess) { ok = 0; cudaGetLastError(); }
		UG_CUDA(ctx, cudaMemcpyAsync(scr, &h, sizeof(h), cudaMemcpyHostToDevice, s));
	}
	static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle travels in 64 bytes");
	UG_NCCL(ctx, N.Broadcast(scr, scr, 64, ncclUint8, 0, comm, s));
	/* 3. the others map it */
	if (rank != 0) {
		UG_CUDA(ctx, cudaMemcpyAsync(&h, scr, sizeof(h), cudaMemcpyDeviceToHost, s));
		UG_CUDA(ctx, cudaStreamSynchronize(s));
		void *p = nullptr;
		if (getenv("UG_NO_PUSH") || cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { ok = 0; cudaGetLastError(); p = nullptr; }
		ctx->peer_full = (unsigned char *)p;
	}
	/* 4. push only if it works everywhere */
	UG_CUDA(ctx, cudaMemcpyAsync(scr + 64, &ok, 4, cudaMemcpyHostToDevice, s));
	UG_NCCL(ctx, N.AllReduce(scr + 64, scr + 64, 1, ncclUint32, ncclMin, comm, s));
	UG_CUDA(ctx, cudaMemcpyAsync(&ok, scr + 64, 4, cudaMemcpyDeviceToHost, s));
	UG_CUDA(ctx, cudaStreamSynchronize(s));
	ctx->use_push = ok != 0;
	if (!ctx->use_push) unmap_peer(ctx);
	ctx->rgb_full_bytes = bytes;
	UG_CUDA(ctx, cudaMemsetAsync(scr + 200, 0, 4, s)); /* this rank's count of gathered frames (flag handshakes) */
	UG_CUDA(ctx, cudaStreamSynchronize(s));
	return UG_OK;
}

/* push transport: "rank 0 is past its read-out of the previous frame".  Issued at the head of the frame on a side
 * branch, so its latency hides behind the frame's own kernels. */
static void *comm_flags(ug_context *ctx)
{
	unsigned char *base = ctx->comm_rank == 0 ? ctx->d_rgb_full : ctx->peer_full;
	return base ? base + ctx->rgb_full_bytes : nullptr;
}
int ugc_ready(ug_context *ctx, cudaStream_t s, bool deferred)
{
	if (!ctx->use_push) return UG_OK;
	unsigned char *scr = (unsigned char *)ctx->d_comm_scratch;
	if (ctx->flag_handshake && !deferred) { /* flags in rank 0's mapped buffer instead of a collective (k_shade.cu) */
		ugk_comm_ready(s, comm_flags(ctx), (u32 *)(scr + 200), ctx->comm_rank);
		return UG_OK;
	}
	UG_NCCL(ctx, N.AllReduce(scr + 128, scr + 128, 1, ncclUint32, ncclSum, (ncclComm_t)ctx->comm, s));
	return UG_OK;
}

/* ns of this rank's own work in the last gathered frame (first kernel -> in front of the "done" handshake); flag handshakes only */
u32 *ugc_own_word(ug_context *ctx)
{
	return (ctx->comm && ctx->use_push && ctx->flag_handshake && ctx->d_comm_scratch) ? (u32 *)((unsigned char *)ctx->d_comm_scratch + 204) : nullptr;
}
u32 *ugc_overflow_word(ug_context *ctx)
{
	return (ctx->comm && ctx->use_push && ctx->d_comm_scratch) ? (u32 *)((unsigned char *)ctx->d_comm_scratch + 136) : nullptr;
}

/* host-side collective decision (e.g. "did any rank overflow its pair arena?"): max over the ranks of *value */
int ugc_allreduce_max(ug_context *ctx, cudaStream_t s, unsigned *value)
{
	if (!ctx->comm || ctx->comm_world <= 1) return UG_OK;
	if (!ctx->d_comm_scratch) {
		UG_CUDA(ctx, cudaMalloc((void **)&ctx->d_comm_scratch, 256));
		UG_CUDA(ctx, cudaMemset(ctx->d_comm_scratch, 0, 256));
	}
	unsigned char *scr = (unsigned char *)ctx->d_comm_scratch;
	UG_CUDA(ctx, cudaMemcpyAsync(scr + 192, value, 4, cudaMemcpyHostToDevice, s));
	UG_NCCL(ctx, N.AllReduce(scr + 192, scr + 192, 1, ncclUint32, ncclMax, (ncclComm_t)ctx->comm, s));
	UG_CUDA(ctx, cudaMemcpyAsync(value, scr + 192, 4, cudaMemcpyDeviceToHost, s));
	UG_CUDA(ctx, cudaStreamSynchronize(s));
	return UG_OK;
}

/* deferred gather: the handshake that makes the last frame whole (what the next frame's "ready" would have been) */
int ugc_flush(ug_context *ctx, cudaStream_t s)
{
	if (!ctx->use_push) return UG_OK; /* ncclRecv completes with the data: nothing is ever pending */
	unsigned char *scr = (unsigned char *)ctx->d_comm_scratch;
	UG_NCCL(ctx, N.AllReduce(scr + 132, scr + 132, 1, ncclUint32, ncclSum, (ncclComm_t)ctx->comm, s));
	return UG_OK;
}

/* strips -> rank 0's full image.  bytes[r] / offs[r] describe every rank's strip inside the full image.  The caller has
 * ordered ugc_ready() before this on the same stream (directly or through an event). */
int ugc_gather_strips(ug_context *ctx, cudaStream_t s, const unsigned char *strip, size_t slot_offset, const size_t *bytes, const size_t *offs, bool deferred,
                      bool already_stored)
{
	ncclComm_t comm = (ncclComm_t)ctx->comm;
	const int rank = ctx->comm_rank, world = ctx->comm_world;
	unsigned char *full = ctx->d_rgb_full ? ctx->d_rgb_full + slot_offset : nullptr;
	if (ctx->use_push) {
		unsigned char *scr = (unsigned char *)ctx->d_comm_scratch;
		if (already_stored) { /* fused gather: the shading kernels wrote the pixels into rank 0's frame themselves */ }
		else if (rank == 0) UG_CUDA(ctx, cudaMemcpyAsync(full + offs[0], strip, bytes[0], cudaMemcpyDeviceToDevice, s));
		else UG_CUDA(ctx, cudaMemcpyAsync(ctx->peer_full + slot_offset + offs[rank], strip, bytes[rank], cudaMemcpyDefault, s)); /* copy engine, over NVLink */
		/* done: every copy precedes its rank's contribution.  Deferred: the NEXT frame's "ready" handshake is issued after
		 * this copy in stream order on every rank, so it doubles as this frame's "done" and nothing waits here */
		if (!deferred) {
			if (ctx->flag_handshake) ugk_comm_done(s, comm_flags(ctx), (const u32 *)(scr + 200), rank, world, (u32 *)(scr + 136), ctx->d_tstart, (u32 *)(scr + 204));
			else UG_NCCL(ctx, N.AllReduce(scr + 132, scr + 132, 2, ncclUint32, ncclSum, comm, s)); /* [1] = ranks whose pair arena overflowed */
		}
		return UG_OK;
	}
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

/* ---- light grids built once per communicator instead of once per rank (strip mode) -------------------------------------
 * Every rank builds the cell rows [nby*q/world, nby*(q+1)/world) of each light grid and stores its lists into the same
 * place of EVERY rank's copy (P2P stores over NVLink, k_lightshare.cu).  For that every rank exports one buffer (flags page,
 * then per light: span[C], offset[C], vals[world * slot]) and maps the buffers of all the others.  Collective. */
static void unmap_light_peers(ug_context *ctx)
{
	for (int r = 0; r < UG_MAX_RANKS; r++) {
		if (ctx->lg_peer[r] && ctx->lg_peer[r] != ctx->lg_local) cudaIpcCloseMemHandle(ctx->lg_peer[r]); /* by pointer, not by rank: ug_comm_init may have renumbered this context */
		ctx->lg_peer[r] = nullptr;
	}
}
void ugc_unshare_light(ug_context *ctx)
{
	unmap_light_peers(ctx);
	if (ctx->lg_local) cudaFree(ctx->lg_local);
	ctx->lg_local = nullptr;
	ctx->lg_bytes = 0;
	ctx->lg_ok = false;
}
int ugc_share_light(ug_context *ctx, size_t bytes)
{
	ncclComm_t comm = (ncclComm_t)ctx->comm;
	cudaStream_t s = ctx->stream;
	const int rank = ctx->comm_rank, world = ctx->comm_world;
	/* nobody keeps an old buffer mapped when its owner frees it */
	UG_CUDA(ctx, cudaStreamSynchronize(s));
	unmap_light_peers(ctx);
	unsigned char *scr = (unsigned char *)ctx->d_comm_scratch;
	UG_NCCL(ctx, N.AllReduce(scr + 128, scr + 128, 1, ncclUint32, ncclSum, comm, s));
	UG_CUDA(ctx, cudaStreamSynchronize(s));
	if (ctx->lg_local) cudaFree(ctx->lg_local);
	ctx->lg_local = nullptr;
	unsigned int ok = 1;
	cudaIpcMemHandle_t mine;
	memset(&mine, 0, sizeof(mine));
	if (cudaMalloc((void **)&ctx->lg_local, bytes) != cudaSuccess) { ok = 0; cudaGetLastError(); ctx->lg_local = nullptr; }
	else {
		UG_CUDA(ctx, cudaMemsetAsync(ctx->lg_local, 0, UG_COMM_FLAGS_BYTES, s)); /* arrival flags: sequence numbers start at 0 */
		if (getenv("UG_NO_PUSH") || cudaIpcGetMemHandle(&mine, ctx->lg_local) != cudaSuccess) { ok = 0; cudaGetLastError(); }
	}
	if (!ctx->d_lg_handles) UG_CUDA(ctx, cudaMalloc((void **)&ctx->d_lg_handles, 64 * (size_t)(UG_MAX_RANKS + 1)));
	unsigned char *hd = (unsigned char *)ctx->d_lg_handles;
	UG_CUDA(ctx, cudaMemcpyAsync(hd + 64 * (size_t)UG_MAX_RANKS, &mine, 64, cudaMemcpyHostToDevice, s));
	UG_NCCL(ctx, N.AllGather(hd + 64 * (size_t)UG_MAX_RANKS, hd, 64, ncclUint8, comm, s));
	std::vector<cudaIpcMemHandle_t> all((size_t)world);
	UG_CUDA(ctx, cudaMemcpyAsync(all.data(), hd, 64 * (size_t)world, cudaMemcpyDeviceToHost, s));
	UG_CUDA(ctx, cudaStreamSynchronize(s));
	for (int r = 0; r < world && ok; r++) {
		if (r == rank) { ctx->lg_peer[r] = ctx->lg_local; continue; }
		void *p = nullptr;
		if (cudaIpcOpenMemHandle(&p, all[(size_t)r], cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { ok = 0; cudaGetLastError(); p = nullptr; }
		ctx->lg_peer[r] = (unsigned char *)p;
	}
	UG_CUDA(ctx, cudaMemcpyAsync(scr + 64, &ok, 4, cudaMemcpyHostToDevice, s));
	UG_NCCL(ctx, N.AllReduce(scr + 64, scr + 64, 1, ncclUint32, ncclMin, comm, s));
	UG_CUDA(ctx, cudaMemcpyAsync(&ok, scr + 64, 4, cudaMemcpyDeviceToHost, s));
	UG_CUDA(ctx, cudaStreamSynchronize(s));
	ctx->lg_ok = ok != 0;
	ctx->lg_bytes = bytes; /* also when it failed: the attempt is not repeated for the same layout */
	if (!ctx->lg_ok) {
		unmap_light_peers(ctx);
		if (ctx->lg_local) cudaFree(ctx->lg_local);
		ctx->lg_local = nullptr;
	}
	return UG_OK;
}
