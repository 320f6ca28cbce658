/* present_gl.cpp -- SURVEY.md 8(f)-4, the CUDA side of the reference's presenter.
 *
 * The reference renders into an OpenGL pixel buffer object: main() creates it and registers it with CUDA
 * (main.cu:343-347, cudaGLRegisterBufferObject), display() maps it as d_image in front of the frame
 * (main.cu:139, cudaGLMapBufferObject), unmaps it behind it (main.cu:262) and draws it as a texture
 * (main.cu:274-295: glTexSubImage2D from the bound PBO + one quad).  The GL calls stay with the host program -
 * it owns the window and the GL context - but everything CUDA has to do with the PBO is here, so that the shading
 * kernels write the image straight into the PBO and no copy stands between the frame and glTexSubImage2D:
 *
 *   ug_gl_register_pbo(ctx, pbo)   once, after glBufferData            (main.cu:347)
 *   ug_gl_map_pbo(ctx)             per frame, before ug_render_frame   (main.cu:139): d_image = the PBO's storage
 *   ug_gl_unmap_pbo(ctx)           per frame, before the GL draw       (main.cu:262): waits for the frame
 *   ug_gl_unregister_pbo(ctx)      at exit                             (cleanup())
 *
 * Written against the graphics-interop API of the runtime (cudaGraphicsGLRegisterBuffer / cudaGraphicsMapResources),
 * which replaced the cudaGL*BufferObject calls of the reference's CUDA 2.x; the prototype is declared here because
 * <cuda_gl_interop.h> wants <GL/gl.h>, which the build image does not have (GLuint is unsigned int).  The calling
 * thread must have the GL context of the PBO current, as in the reference.  The frame's CUDA graph is captured
 * with the mapped address (frame_key mixes d_rgb in); drivers hand out the same address at every map of an
 * unchanged buffer, so the graph is captured once - if the address ever moves, the next frame re-captures.
 * No GL exists on the B200 image, so this file is compiled and its argument checks are tested, but it has not been
 * run against a GL context.
 */
#include "ugrt_internal.h"

extern "C" cudaError_t cudaGraphicsGLRegisterBuffer(struct cudaGraphicsResource **resource, unsigned int buffer, unsigned int flags);

extern "C" int ug_gl_register_pbo(ug_context *ctx, unsigned int pbo)
{
	if (!ctx) return UG_ERR_ARG;
	if (ctx->gl_mapped) return ug_fail(ctx, UG_ERR_ARG, "ug_gl_register_pbo: the current PBO is still mapped");
	UG_CUDA(ctx, cudaSetDevice(ctx->device));
	if (ctx->gl_resource) {
		UG_CUDA(ctx, cudaGraphicsUnregisterResource((cudaGraphicsResource_t)ctx->gl_resource));
		ctx->gl_resource = nullptr;
	}
	cudaGraphicsResource_t res = nullptr;
	/* the frame rewrites every pixel and never reads the old image: write-discard lets the driver skip preserving it */
	UG_CUDA(ctx, cudaGraphicsGLRegisterBuffer(&res, pbo, cudaGraphicsRegisterFlagsWriteDiscard));
	ctx->gl_resource = res;
	return UG_OK;
}

extern "C" int ug_gl_map_pbo(ug_context *ctx)
{
	if (!ctx) return UG_ERR_ARG;
	if (!ctx->gl_resource) return ug_fail(ctx, UG_ERR_ARG, "ug_gl_map_pbo: no PBO registered");
	if (ctx->gl_mapped) return UG_OK;
	cudaGraphicsResource_t res = (cudaGraphicsResource_t)ctx->gl_resource;
	UG_CUDA(ctx, cudaGraphicsMapResources(1, &res, ctx->stream)); /* GL work on the buffer is ordered in front of the stream */
	void *ptr = nullptr;
	size_t bytes = 0;
	cudaError_t e = cudaGraphicsResourceGetMappedPointer(&ptr, &bytes, res);
	if (e != cudaSuccess || !ptr) {
		cudaGraphicsUnmapResources(1, &res, ctx->stream);
		return ug_fail(ctx, UG_ERR_CUDA, "ug_gl_map_pbo: %s", cudaGetErrorString(e));
	}
	ctx->gl_mapped = true;
	if (ctx->d_rgb_ext != (unsigned char *)ptr || ctx->rgb_ext_bytes != bytes) { /* ug_set_rgb_buffer without its synchronise */
		ctx->d_rgb_ext = (unsigned char *)ptr;
		ctx->rgb_ext_bytes = bytes;
		ug_drop_graphs(ctx);
	}
	return UG_OK; /* a PBO smaller than 3*W*H bytes is refused by the next frame (ensure_buffers), like any external buffer */
}

extern "C" int ug_gl_unmap_pbo(ug_context *ctx)
{
	if (!ctx) return UG_ERR_ARG;
	if (!ctx->gl_mapped) return UG_OK;
	cudaGraphicsResource_t res = (cudaGraphicsResource_t)ctx->gl_resource;
	/* the unmap is ordered behind the frame on the context's stream; GL commands issued after it see the image */
	UG_CUDA(ctx, cudaGraphicsUnmapResources(1, &res, ctx->stream));
	ctx->gl_mapped = false;
	return UG_OK;
}

extern "C" int ug_gl_unregister_pbo(ug_context *ctx)
{
	if (!ctx) return UG_ERR_ARG;
	if (!ctx->gl_resource) return UG_OK;
	int rc = ug_gl_unmap_pbo(ctx);
	if (rc) return rc;
	UG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
	UG_CUDA(ctx, cudaGraphicsUnregisterResource((cudaGraphicsResource_t)ctx->gl_resource));
	ctx->gl_resource = nullptr;
	ctx->d_rgb_ext = nullptr; /* back to the context's own image */
	ctx->rgb_ext_bytes = 0;
	ctx->d_rgb = ctx->d_rgb_own;
	ug_drop_graphs(ctx);
	return UG_OK;
}
