/* dda_sim.c -- development tool (not product, not oracle): counts the work the reflection DDA does per ray on the
 * CPU for a given grid resolution, so the grid parameters can be chosen without GPU time.
 * gcc -O2 -shared -fPIC -o build/libddasim.so tests/devtools/dda_sim.c -lm */
#include <math.h>
#include <stdlib.h>
#include <string.h>

int g_cull = 0;
/* empty-space skipping study: for coarse blocks of 2^k cells per axis (k = 1,2,3), per ray the fine steps taken inside blocks that
 * hold no triangle at all (skippable) and the number of such blocks entered (what a two-level walk would pay instead) */
long g_skip_steps[4], g_skip_blocks[4];
typedef struct { float lo[3], h[3], inv_h[3]; int n[3]; } G;
static int clampi(int v, int a, int b) { return v < a ? a : (v > b ? b : v); }
static int ucell(const G *u, int k, float x) { return clampi((int)floorf((x - u->lo[k]) * u->inv_h[k]), 0, u->n[k] - 1); }

/* verts [V*3], tris [T*3], rays O,r [R*3]; n[3] explicit resolution; out: steps[R], tests[R]; returns pairs */
long dda_sim(const float *verts, int V, const int *tris, int T, const float *O, const float *r, int R, const int *n, int *steps, int *tests, int *nonempty)
{
	G u;
	for (int k = 0; k < 3; k++) {
		float lo = 3e38f, hi = -3e38f;
		for (int v = 0; v < V; v++) { float x = verts[3 * v + k]; if (x < lo) lo = x; if (x > hi) hi = x; }
		float pad = 1e-3f * fmaxf(hi - lo, 1e-3f);
		lo -= pad;
		float ext = (hi + pad) - lo;
		u.lo[k] = lo; u.n[k] = n[k]; u.h[k] = ext / n[k]; u.inv_h[k] = n[k] / ext;
	}
	long C = (long)n[0] * n[1] * n[2];
	int *span = calloc(C + 1, sizeof(int));
	int *rect = malloc(sizeof(int) * 6 * T);
	long P = 0;
	for (int t = 0; t < T; t++) {
		for (int k = 0; k < 3; k++) {
			float a = verts[3 * tris[3 * t] + k], b = verts[3 * tris[3 * t + 1] + k], c = verts[3 * tris[3 * t + 2] + k];
			float mn = fminf(a, fminf(b, c)), mx = fmaxf(a, fmaxf(b, c));
			float eps = 1e-4f * u.h[k] + 1e-6f * fmaxf(fabsf(mn), fabsf(mx));
			rect[6 * t + 2 * k] = ucell(&u, k, mn - eps); rect[6 * t + 2 * k + 1] = ucell(&u, k, mx + eps);
		}
		for (int x = rect[6 * t]; x <= rect[6 * t + 1]; x++) for (int y = rect[6 * t + 2]; y <= rect[6 * t + 3]; y++) for (int z = rect[6 * t + 4]; z <= rect[6 * t + 5]; z++) {
			span[((long)x * n[1] + y) * n[2] + z]++; P++;
		}
	}
	long *off = malloc(sizeof(long) * (C + 1));
	off[0] = 0;
	for (long c = 0; c < C; c++) off[c + 1] = off[c] + span[c];
	int *vals = malloc(sizeof(int) * (P + 1));
	int *fill = calloc(C, sizeof(int));
	for (int t = 0; t < T; t++)
		for (int x = rect[6 * t]; x <= rect[6 * t + 1]; x++) for (int y = rect[6 * t + 2]; y <= rect[6 * t + 3]; y++) for (int z = rect[6 * t + 4]; z <= rect[6 * t + 5]; z++) {
			long c = ((long)x * n[1] + y) * n[2] + z; vals[off[c] + fill[c]++] = t;
		}
	float *ubox = malloc(sizeof(float) * 6 * C);
	for (long c = 0; c < C; c++) {
		long cz = c % n[2], cy = (c / n[2]) % n[1], cx = c / ((long)n[1] * n[2]);
		long cc[3] = {cx, cy, cz};
		for (int k = 0; k < 3; k++) { ubox[6 * c + k] = 3e38f; ubox[6 * c + 3 + k] = -3e38f; }
		for (int q = 0; q < span[c]; q++) {
			int t = vals[off[c] + q];
			for (int k = 0; k < 3; k++) {
				float a = verts[3 * tris[3 * t] + k], b = verts[3 * tris[3 * t + 1] + k], c2 = verts[3 * tris[3 * t + 2] + k];
				float mn = fminf(a, fminf(b, c2)), mx = fmaxf(a, fmaxf(b, c2));
				float clo = u.lo[k] + u.h[k] * cc[k], chi = clo + u.h[k];
				mn = fmaxf(mn, clo); mx = fminf(mx, chi);
				/* quantize outward to 1/1024 of the cell */
				float ql = floorf((mn - clo) * u.inv_h[k] * 1023.0f), qh = ceilf((mx - clo) * u.inv_h[k] * 1023.0f);
				mn = clo + ql / 1023.0f * u.h[k]; mx = clo + qh / 1023.0f * u.h[k];
				if (mn < ubox[6 * c + k]) ubox[6 * c + k] = mn;
				if (mx > ubox[6 * c + 3 + k]) ubox[6 * c + 3 + k] = mx;
			}
		}
	}
	unsigned char *occ[4] = {0, 0, 0, 0};
	int cn[4][3];
	for (int k = 1; k <= 3; k++) {
		for (int a = 0; a < 3; a++) cn[k][a] = (n[a] + (1 << k) - 1) >> k;
		occ[k] = calloc((size_t)cn[k][0] * cn[k][1] * cn[k][2], 1);
		for (long c = 0; c < C; c++) if (span[c]) {
			long cz = c % n[2], cy = (c / n[2]) % n[1], cx = c / ((long)n[1] * n[2]);
			occ[k][((cx >> k) * cn[k][1] + (cy >> k)) * cn[k][2] + (cz >> k)] = 1;
		}
		g_skip_steps[k] = g_skip_blocks[k] = 0;
	}
	float hmin = fminf(u.h[0], fminf(u.h[1], u.h[2]));
	for (int i = 0; i < R; i++) {
		const float *o = O + 3 * i, *d = r + 3 * i;
		float t_in = 0, t_out = 3e38f; int miss = 0;
		for (int k = 0; k < 3; k++) {
			float hi = u.lo[k] + u.h[k] * n[k];
			if (d[k] != 0) { float t0 = (u.lo[k] - o[k]) / d[k], t1 = (hi - o[k]) / d[k]; if (t0 > t1) { float s = t0; t0 = t1; t1 = s; } t_in = fmaxf(t_in, t0); t_out = fminf(t_out, t1); }
			else if (o[k] < u.lo[k] || o[k] > hi) miss = 1;
		}
		int ns = 0, nt = 0, ne = 0, nc = 0;
		if (!miss && t_in <= t_out) {
			int ci[3], st[3]; float tmax[3], tdel[3];
			for (int k = 0; k < 3; k++) {
				float s = o[k] + t_in * d[k];
				ci[k] = ucell(&u, k, s);
				st[k] = d[k] > 0 ? 1 : (d[k] < 0 ? -1 : 0);
				if (st[k]) { float bound = u.lo[k] + u.h[k] * (ci[k] + (st[k] > 0)); tmax[k] = (bound - o[k]) / d[k]; tdel[k] = u.h[k] / fabsf(d[k]); }
				else tmax[k] = tdel[k] = 3e38f;
			}
			float best = 99999999.9f; int hit = -1;
			float t_prev = t_in;
			long last_block[4] = {-1, -1, -1, -1};
			for (;;) {
				long c = ((long)ci[0] * n[1] + ci[1]) * n[2] + ci[2];
				for (int k = 1; k <= 3; k++) {
					long b = (((long)(ci[0] >> k)) * cn[k][1] + (ci[1] >> k)) * cn[k][2] + (ci[2] >> k);
					if (!occ[k][b]) { g_skip_steps[k]++; if (b != last_block[k]) g_skip_blocks[k]++; }
					last_block[k] = b;
				}
				int ax = (tmax[0] <= tmax[1]) ? ((tmax[0] <= tmax[2]) ? 0 : 2) : ((tmax[1] <= tmax[2]) ? 1 : 2);
				float texit = tmax[ax];
				ns++;
				float slo[3], shi[3];
				for (int k = 0; k < 3; k++) {
					float a = o[k] + t_prev * d[k], b = o[k] + texit * d[k];
					float e = 1e-4f * u.h[k] + 1e-6f * fmaxf(fabsf(a), fabsf(b));
					slo[k] = fminf(a, b) - e; shi[k] = fmaxf(a, b) + e;
				}
				t_prev = texit;
				int cellout = 0;
				if (g_cull >= 2) for (int k = 0; k < 3; k++) if (ubox[6 * c + k] > shi[k] || ubox[6 * c + 3 + k] < slo[k]) cellout = 1;
				if (span[c]) nc++;
				for (int q = 0; q < (cellout ? 0 : span[c]); q++) {
					int t = vals[off[c] + q];
					if (g_cull) {
						int out = 0;
						for (int k = 0; k < 3; k++) {
							float a = verts[3 * tris[3 * t] + k], b = verts[3 * tris[3 * t + 1] + k], cc = verts[3 * tris[3 * t + 2] + k];
							float mn = fminf(a, fminf(b, cc)), mx = fmaxf(a, fmaxf(b, cc));
							if (mn > shi[k] || mx < slo[k]) out = 1;
						}
						ne++;
						if (out) continue;
					}
					nt++;
					const float *v0 = verts + 3 * tris[3 * t], *v1 = verts + 3 * tris[3 * t + 1], *v2 = verts + 3 * tris[3 * t + 2];
					float e1[3] = {v1[0] - v0[0], v1[1] - v0[1], v1[2] - v0[2]}, e2[3] = {v2[0] - v0[0], v2[1] - v0[1], v2[2] - v0[2]};
					float p[3] = {d[1] * e2[2] - d[2] * e2[1], d[2] * e2[0] - d[0] * e2[2], d[0] * e2[1] - d[1] * e2[0]};
					float det = e1[0] * p[0] + e1[1] * p[1] + e1[2] * p[2];
					if (fabsf(det) <= 1e-21f) continue;
					float inv = 1.0f / det, tv[3] = {o[0] - v0[0], o[1] - v0[1], o[2] - v0[2]};
					float uu = (tv[0] * p[0] + tv[1] * p[1] + tv[2] * p[2]) * inv;
					if (uu < 0 || uu > 1) continue;
					float qv[3] = {tv[1] * e1[2] - tv[2] * e1[1], tv[2] * e1[0] - tv[0] * e1[2], tv[0] * e1[1] - tv[1] * e1[0]};
					float vv = (d[0] * qv[0] + d[1] * qv[1] + d[2] * qv[2]) * inv;
					if (vv < 0 || uu + vv > 1) continue;
					float tt = (e2[0] * qv[0] + e2[1] * qv[1] + e2[2] * qv[2]) * inv;
					if (tt > 0 && (tt < best || (tt == best && t < hit))) { best = tt; hit = t; }
				}
				if (hit >= 0 && best < texit * 0.999f - 1e-4f * hmin) break;
				ci[ax] += st[ax];
				if (st[ax] == 0 || ci[ax] < 0 || ci[ax] >= n[ax]) break;
				tmax[ax] += tdel[ax];
			}
		}
		steps[i] = ns | (nc << 16); tests[i] = nt; nonempty[i] = ne;
	}
	for (int k = 1; k <= 3; k++) free(occ[k]);
	free(ubox); free(span); free(rect); free(off); free(vals); free(fill);
	return P;
}
