// plugin_shim.cpp — a Linux implementation of the reference's `core/plugin.h` interface
// (diverse/diverse_base/source/core/plugin.h: Plugin, PluginManager), written here because the reference's
// own plugin.cpp does not compile with g++ 13 (plugin.cpp:93,125 pass a runtime string to std::format).
// It is linked ONLY into the test build of the unmodified diverseshot-cli (divshot_b200/build.py:
// build_reference_cli) so that the CLI can dlopen our libgstrain.so exactly as plugin.cpp:74,89,153-166 would:
// dlopen("lib<name>.so", RTLD_LAZY | RTLD_LOCAL) through the normal loader search path, dlsym per symbol.
#include <core/plugin.h>

#include <cstdio>

namespace diverse {

Plugin::Plugin(const std::string& shortName, const Path& path)
    : m_createInstance(nullptr), m_DescriptInstance(nullptr), m_shortName(shortName), m_path(path), m_handle(nullptr) {
    const std::string file = (path.has_parent_path() ? path.parent_path().string() + "/" : std::string()) + "lib" +
                             shortName + ".so";
    m_handle = dlopen(file.c_str(), RTLD_LAZY | RTLD_LOCAL);
    if (!m_handle) {
        std::fprintf(stderr, "plugin_shim: dlopen(%s) failed: %s\n", file.c_str(), dlerror());
        return;
    }
    m_DescriptInstance = reinterpret_cast<getDescriptionFunc>(dlsym(m_handle, "get_description"));
    m_createInstance = reinterpret_cast<createInstanceFunc>(dlsym(m_handle, "create_instance"));
}
Plugin::~Plugin() {
    if (m_handle) dlclose(m_handle);
}
ObjHandle Plugin::create_instance() const { return m_createInstance ? m_createInstance() : nullptr; }
std::string Plugin::get_description() const { return m_DescriptInstance ? m_DescriptInstance() : ""; }
const Path& Plugin::path() const { return m_path; }
const std::string& Plugin::short_name() const { return m_shortName; }
bool Plugin::has_symbol(const std::string& sym) const { return m_handle && dlsym(m_handle, sym.c_str()) != nullptr; }
void* Plugin::get_symbol(const std::string& sym) { return dlsym(m_handle, sym.c_str()); }
auto Plugin::get_last_error_as_string() const -> std::string {
    const char* e = dlerror();
    return e ? e : "";
}

ObjHandle PluginManager::create_object(const std::string& name) {
    ensure_plugin_loaded(name);
    return m_plugins[name] ? m_plugins[name]->create_instance() : nullptr;
}
std::vector<std::string> PluginManager::get_loaded_plugins() {
    std::vector<std::string> l;
    for (auto& kv : m_plugins) l.push_back(kv.first);
    return l;
}
Plugin* PluginManager::get_plugin(const std::string& name) {
    auto it = m_plugins.find(name);
    return it == m_plugins.end() ? nullptr : it->second.get();
}
bool PluginManager::ensure_plugin_loaded(const std::string& name) {
    if (m_plugins[name]) return true;
    Path p(name);
    m_plugins[name] = std::make_unique<Plugin>(p.filename().string(), p);
    std::printf("Successfully loaded plugin %s\n", name.c_str());
    return true;
}

}  // namespace diverse
