// PETSc-free options store with the reference's accessor names (src/common_utils.cpp:67-133).
#include <fstream>
#include <map>
#include <sstream>
#include "common_utils.hpp"
#include "s3d_device.hpp"

namespace {
std::map<std::string, std::string> &db()
{
	static std::map<std::string, std::string> d;
	return d;
}

bool is_key(const std::string &t)
{
	return t.size() >= 2 && t[0] == '-' && !(t[1] >= '0' && t[1] <= '9') && t[1] != '.';
}

void absorb(const std::vector<std::string> &tok)
{
	for (size_t i = 0; i < tok.size(); i++) {
		if (!is_key(tok[i])) continue;
		std::string val;
		if (i + 1 < tok.size() && !is_key(tok[i + 1])) val = tok[i + 1];
		db()[tok[i]] = val;
	}
}

const std::string &need(const std::string &key)
{
	auto it = db().find(key);
	if (it == db().end()) throw NonExistentPetscOpion(std::string("Could not find ") + key);
	return it->second;
}
}  // namespace

NonExistentPetscOpion::NonExistentPetscOpion(const std::string &msg) : std::runtime_error(msg) {}

namespace s3d {

void options_load_file(const std::string &path)
{
	std::ifstream f(path);
	if (!f) throw std::runtime_error("Could not open options file " + path);
	std::vector<std::string> tok;
	std::string line;
	while (std::getline(f, line)) {
		const size_t hash = line.find('#');
		if (hash != std::string::npos) line.erase(hash);
		std::istringstream ss(line);
		std::string t;
		while (ss >> t) tok.push_back(t);
	}
	absorb(tok);
}

void options_from_args(int argc, char **argv)
{
	std::vector<std::string> tok;
	for (int i = 1; i < argc; i++) tok.push_back(argv[i]);
	for (size_t i = 0; i + 1 < tok.size(); i++)
		if (tok[i] == "-options_file") options_load_file(tok[i + 1]);
	absorb(tok);
}

void options_set(const std::string &key, const std::string &value) { db()[key] = value; }
void options_clear() { db().clear(); }
bool options_has(const std::string &key) { return db().count(key) != 0; }

}  // namespace s3d

std::string petscoptions_get_string(const std::string optionname, const size_t sz)
{
	std::string v = need(optionname);
	if (sz > 0 && v.size() >= sz) v.resize(sz - 1);
	return v;
}

sreal petscoptions_get_real(const std::string optionname) { return std::strtod(need(optionname).c_str(), nullptr); }

int petscoptions_get_int(const std::string optionname) { return (int)std::strtol(need(optionname).c_str(), nullptr, 10); }

bool petscoptions_get_bool(const std::string optionname)
{
	const std::string &v = need(optionname);
	return v.empty() || v == "true" || v == "1" || v == "yes" || v == "on" || v == "TRUE";
}

std::vector<int> petscoptions_get_array_int(const std::string optionname, const int maxlen)
{
	auto it = db().find(optionname);
	if (it == db().end()) throw NonExistentPetscOpion(std::string("Array ") + optionname + std::string(" not set"));
	std::vector<int> arr;
	std::stringstream ss(it->second);
	std::string item;
	while (std::getline(ss, item, ',') && (int)arr.size() < maxlen) arr.push_back((int)std::strtol(item.c_str(), nullptr, 10));
	return arr;
}

int get_mpi_rank(int) { return s3d::Device::get().rank(); }
int get_mpi_size(int) { return s3d::Device::get().nranks(); }
