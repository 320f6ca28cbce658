/* obj_loader.cpp -- the loader boundary of scene.h (`Model::load_model` static branch :225-331,
 * `Model::some_material` :370-439) over an OBJ/MTL reader that keeps obj_parser/'s observable
 * semantics for the fields the renderer uses (obj_parser/obj_parser.cpp:296-421, :177-294;
 * obj_parser/list.cpp:115-126):
 *   - lines are read in 500-byte pieces and tokenised on " \t\n\r"; a first token starting with '#'
 *     is a comment;
 *   - `v x y z` is parsed with atof (double) and narrowed to float by load_model;
 *   - `f a b c ...` keeps the vertex index before the first '/', converts it like
 *     obj_convert_to_list_index (0 -> -1, negative -> relative, else index-1) and load_model keeps the
 *     first three corners only (quads are truncated, scene.h:251-253);
 *   - `mtllib file` is opened relative to the current directory and only its `newmtl` names matter
 *     here; a name keeps whatever followed it on the line (the MTL reader splits on " \t" only);
 *   - `usemtl name` selects the FIRST material whose stored name starts with `name` (prefix match),
 *     -1 if none; faces before any usemtl get -1.
 * In a drop-in build the reference's own obj_parser/ stays in place and Model hands its arrays to
 * ug_model_set(); this reader lets the library (and the tests) stand alone.
 */
#include "ugrt_internal.h"
#include <stdlib.h>
#include <string.h>

namespace {

const char *kSpace = " \t\n\r";

struct MaterialNames { std::vector<std::string> names; };

void read_material_names(const char *file, MaterialNames &out)
{
	FILE *fp = fopen(file, "r");
	if (!fp) {
		fprintf(stderr, "Error reading file: %s\n", file);
		return;
	}
	char line[500];
	while (fgets(line, sizeof(line), fp)) {
		char *tok = strtok(line, kSpace);
		if (tok && strcmp(tok, "newmtl") == 0) {
			char *name = strtok(NULL, " \t");
			if (name) out.names.push_back(std::string(name).substr(0, 254));
		}
	}
	fclose(fp);
}

int lookup(const MaterialNames &m, const char *want)
{
	const size_t n = strlen(want);
	for (size_t i = 0; i < m.names.size(); i++)
		if (m.names[i].compare(0, n, want) == 0 && m.names[i].size() >= n) return (int)i;
	return -1;
}

} // namespace

/* vertices only: the per-frame keyframe read of Model::tmp_model (scene.h:86-111 uses nothing else of the file) */
extern "C" int ug_obj_parse_vertices(const char *obj_path, int *num_vertices, float *vertexlist, int capacity)
{
	if (!obj_path || !num_vertices) return UG_ERR_ARG;
	FILE *fp = fopen(obj_path, "r");
	if (!fp) return UG_ERR_IO;
	char line[500];
	int nv = 0;
	while (fgets(line, sizeof(line), fp)) {
		char *tok = strtok(line, kSpace);
		if (!tok || strcmp(tok, "v") != 0) continue;
		for (int k = 0; k < 3; k++) {
			const char *s = strtok(NULL, kSpace);
			const double e = s ? atof(s) : 0.0;
			if (vertexlist && nv < capacity) vertexlist[(size_t)3 * nv + k] = (float)e;
		}
		nv++;
	}
	fclose(fp);
	*num_vertices = nv;
	return UG_OK;
}

extern "C" int ug_obj_parse(const char *obj_path, int *num_vertices, int *num_faces, float *vertexlist, int *facelist, int *matidx)
{
	if (!obj_path || !num_vertices || !num_faces) return UG_ERR_ARG;
	FILE *fp = fopen(obj_path, "r");
	if (!fp) return UG_ERR_IO;
	const bool store = vertexlist && facelist && matidx;
	MaterialNames mats;
	char line[500];
	int nv = 0, nf = 0, current = -1;
	while (fgets(line, sizeof(line), fp)) {
		char *tok = strtok(line, kSpace);
		if (!tok || tok[0] == '#') continue;
		if (strcmp(tok, "v") == 0) {
			for (int k = 0; k < 3; k++) {
				const char *s = strtok(NULL, kSpace);
				const double e = s ? atof(s) : 0.0;
				if (store) vertexlist[(size_t)3 * nv + k] = (float)e;
			}
			nv++;
		} else if (strcmp(tok, "f") == 0) {
			int corner[3] = {0, 0, 0}, n = 0;
			for (char *s = strtok(NULL, kSpace); s; s = strtok(NULL, kSpace), n++)
				if (n < 3) corner[n] = atoi(s);
			if (store) {
				for (int k = 0; k < 3; k++) {
					const int idx = corner[k];
					facelist[(size_t)3 * nf + k] = idx == 0 ? -1 : (idx < 0 ? nv + idx : idx - 1);
				}
				matidx[nf] = current;
			}
			nf++;
		} else if (strcmp(tok, "usemtl") == 0) {
			const char *s = strtok(NULL, kSpace);
			current = s ? lookup(mats, s) : -1;
		} else if (strcmp(tok, "mtllib") == 0) {
			const char *s = strtok(NULL, kSpace);
			if (s) {
				const std::string file(s);
				read_material_names(file.c_str(), mats);
			}
		}
	}
	fclose(fp);
	*num_vertices = nv;
	*num_faces = nf;
	return UG_OK;
}

/* scene.h:370-439: count `newmtl`, then per material: 3 words, 3 floats (ambient), 1 word, 3 floats
 * (diffuse), 11 words, 1 word (texture file or NA) */
extern "C" int ug_material_parse(const char *file, int *num_materials, float *materiallist)
{
	if (!file || !num_materials) return UG_ERR_ARG;
	FILE *fp = fopen(file, "r");
	if (!fp) return UG_ERR_IO;
	char word[512];
	int count = 0;
	while (fscanf(fp, "%511s", word) != EOF)
		if (strcmp(word, "newmtl") == 0) count++;
	*num_materials = count;
	if (materiallist) {
		rewind(fp);
		for (int m = 0; m < count; m++) {
			float *dst = materiallist + (size_t)UG_MATERIAL_SIZE * m;
			for (int i = 0; i < 3; i++) if (fscanf(fp, "%511s", word) != 1) break;
			for (int i = 0; i < 3; i++) if (fscanf(fp, "%f", dst + i) != 1) break;
			if (fscanf(fp, "%511s", word) != 1) break;
			for (int i = 0; i < 3; i++) if (fscanf(fp, "%f", dst + 3 + i) != 1) break;
			for (int i = 0; i < 12; i++) if (fscanf(fp, "%511s", word) != 1) break;
		}
	}
	fclose(fp);
	return UG_OK;
}
