/* TEST INFRASTRUCTURE ONLY (oracle build).
 * Bodies for the PETSc symbols declared in the stub headers: a std::map options database
 * (with "-options_file x.perc" parsing: one "-key value" per line, '#' starts a comment),
 * and aborting bodies for the DM/Vec/Mat calls, which the native path never reaches.
 */
#include <map>
#include <fstream>
#include <sstream>
#include <vector>
#include "petscdm.h"

static std::map<std::string, std::string> &db()
{
	static std::map<std::string, std::string> d;
	return d;
}

void s3d_stub_options_clear() { db().clear(); }

static void parse_tokens(const std::vector<std::string> &tok)
{
	for (size_t i = 0; i < tok.size(); i++) {
		if (tok[i].size() < 2 || tok[i][0] != '-' || (tok[i][1] >= '0' && tok[i][1] <= '9'))
			continue;
		std::string val;
		if (i + 1 < tok.size()) {
			const std::string &nx = tok[i + 1];
			const bool nextIsKey = nx.size() >= 2 && nx[0] == '-' && !(nx[1] >= '0' && nx[1] <= '9') && nx[1] != '.';
			if (!nextIsKey) val = nx;
		}
		db()[tok[i]] = val;
	}
}

static void parse_file(const char *path)
{
	std::ifstream f(path);
	if (!f) throw std::runtime_error(std::string("stub: cannot open options file ") + path);
	std::string line;
	std::vector<std::string> tok;
	while (std::getline(f, line)) {
		const size_t h = line.find('#');
		if (h != std::string::npos) line.erase(h);
		std::istringstream ss(line);
		std::string t;
		while (ss >> t) tok.push_back(t);
	}
	parse_tokens(tok);
}

PetscErrorCode PetscInitialize(int *argc, char ***argv, const char *, const char *)
{
	std::vector<std::string> tok;
	for (int i = 1; argc && i < *argc; i++) tok.push_back((*argv)[i]);
	/* files first so the command line overrides them, as PETSc does */
	for (size_t i = 0; i + 1 < tok.size(); i++)
		if (tok[i] == "-options_file") parse_file(tok[i + 1].c_str());
	parse_tokens(tok);
	return 0;
}
PetscErrorCode PetscFinalize() { return 0; }

static bool lookup(const char *name, std::string &v)
{
	auto it = db().find(name);
	if (it == db().end()) return false;
	v = it->second;
	return true;
}

PetscErrorCode PetscOptionsGetString(PetscOptions, const char *, const char *name, char *out, size_t len,
                                     PetscBool *set)
{
	std::string v;
	*set = lookup(name, v) ? PETSC_TRUE : PETSC_FALSE;
	if (*set) { std::strncpy(out, v.c_str(), len); out[len - 1] = 0; }
	return 0;
}
PetscErrorCode PetscOptionsGetReal(PetscOptions, const char *, const char *name, PetscReal *out, PetscBool *set)
{
	std::string v;
	*set = lookup(name, v) ? PETSC_TRUE : PETSC_FALSE;
	if (*set) *out = std::strtod(v.c_str(), nullptr);
	return 0;
}
PetscErrorCode PetscOptionsGetInt(PetscOptions, const char *, const char *name, PetscInt *out, PetscBool *set)
{
	std::string v;
	*set = lookup(name, v) ? PETSC_TRUE : PETSC_FALSE;
	if (*set) *out = (PetscInt)std::strtol(v.c_str(), nullptr, 10);
	return 0;
}
PetscErrorCode PetscOptionsGetBool(PetscOptions, const char *, const char *name, PetscBool *out, PetscBool *set)
{
	std::string v;
	*set = lookup(name, v) ? PETSC_TRUE : PETSC_FALSE;
	if (*set)
		*out = (v.empty() || v == "true" || v == "1" || v == "yes" || v == "on" || v == "TRUE")
		           ? PETSC_TRUE : PETSC_FALSE;
	return 0;
}
PetscErrorCode PetscOptionsGetIntArray(PetscOptions, const char *, const char *name, PetscInt *out,
                                       PetscInt *n, PetscBool *set)
{
	std::string v;
	*set = lookup(name, v) ? PETSC_TRUE : PETSC_FALSE;
	if (!*set) { *n = 0; return 0; }
	int cnt = 0;
	std::stringstream ss(v);
	std::string item;
	while (std::getline(ss, item, ',') && cnt < *n) out[cnt++] = (PetscInt)std::strtol(item.c_str(), nullptr, 10);
	*n = cnt;
	return 0;
}
PetscErrorCode PetscOptionsSetValue(PetscOptions, const char *name, const char *value)
{
	db()[name] = value ? value : "";
	return 0;
}

#define S3D_NO_PETSC(fn) { std::fprintf(stderr, "oracle stub: " fn " needs real PETSc\n"); std::abort(); }
PetscErrorCode DMDACreate3d(MPI_Comm, DMBoundaryType, DMBoundaryType, DMBoundaryType, DMDAStencilType,
                            PetscInt, PetscInt, PetscInt, PetscInt, PetscInt, PetscInt, PetscInt,
                            PetscInt, const PetscInt *, const PetscInt *, const PetscInt *, DM *)
S3D_NO_PETSC("DMDACreate3d")
PetscErrorCode DMSetUp(DM) S3D_NO_PETSC("DMSetUp")
PetscErrorCode DMDAGetInfo(DM, PetscInt *, PetscInt *, PetscInt *, PetscInt *, PetscInt *, PetscInt *,
                           PetscInt *, PetscInt *, PetscInt *, DMBoundaryType *, DMBoundaryType *,
                           DMBoundaryType *, DMDAStencilType *) S3D_NO_PETSC("DMDAGetInfo")
PetscErrorCode DMDAGetCorners(DM, PetscInt *, PetscInt *, PetscInt *, PetscInt *, PetscInt *, PetscInt *)
S3D_NO_PETSC("DMDAGetCorners")
PetscErrorCode DMDAVecGetArray(DM, Vec, void *) S3D_NO_PETSC("DMDAVecGetArray")
PetscErrorCode DMDAVecRestoreArray(DM, Vec, void *) S3D_NO_PETSC("DMDAVecRestoreArray")
PetscErrorCode MatSetValuesStencil(Mat, PetscInt, const MatStencil *, PetscInt, const MatStencil *,
                                   const PetscScalar *, InsertMode) S3D_NO_PETSC("MatSetValuesStencil")
PetscErrorCode VecDuplicate(Vec, Vec *) S3D_NO_PETSC("VecDuplicate")
PetscErrorCode VecCopy(Vec, Vec) S3D_NO_PETSC("VecCopy")
PetscErrorCode VecAXPY(Vec, PetscScalar, Vec) S3D_NO_PETSC("VecAXPY")
PetscErrorCode VecDestroy(Vec *) S3D_NO_PETSC("VecDestroy")
