// Ordered-map backing for oracle/shim/Judy.h (test infrastructure, NOT product code).
// Semantics follow the published Judy API: JudySL keys are NUL-terminated strings compared as
// unsigned bytes; First/Next/Last/Prev write the found key back into the caller's buffer.
#include "Judy.h"
#include <cstring>
#include <map>
#include <string>

namespace {
typedef std::map<std::string, Word_t> SLMap;  // std::string orders by unsigned byte (char_traits<char>)
typedef std::map<Word_t, Word_t> LMap;
inline SLMap* sl(Pcvoid_t a) { return (SLMap*)a; }
inline LMap* lm(Pcvoid_t a) { return (LMap*)a; }
inline PPvoid_t found(SLMap::iterator it, uint8_t* key) {
    std::memcpy(key, it->first.c_str(), it->first.size() + 1);
    return (PPvoid_t)&it->second;
}
}  // namespace

extern "C" {

PPvoid_t JudySLIns(PPvoid_t arr, const uint8_t* key, void*) {
    if (*arr == NULL) *arr = new SLMap();
    return (PPvoid_t) & (*sl(*arr))[std::string((const char*)key)];
}
PPvoid_t JudySLGet(Pcvoid_t arr, const uint8_t* key, void*) {
    if (!arr) return NULL;
    SLMap::iterator it = sl(arr)->find(std::string((const char*)key));
    return it == sl(arr)->end() ? NULL : (PPvoid_t)&it->second;
}
PPvoid_t JudySLFirst(Pcvoid_t arr, uint8_t* key, void*) {  // first key >= *key
    if (!arr) return NULL;
    SLMap::iterator it = sl(arr)->lower_bound(std::string((const char*)key));
    return it == sl(arr)->end() ? NULL : found(it, key);
}
PPvoid_t JudySLNext(Pcvoid_t arr, uint8_t* key, void*) {  // first key > *key
    if (!arr) return NULL;
    SLMap::iterator it = sl(arr)->upper_bound(std::string((const char*)key));
    return it == sl(arr)->end() ? NULL : found(it, key);
}
PPvoid_t JudySLLast(Pcvoid_t arr, uint8_t* key, void*) {  // last key <= *key
    if (!arr) return NULL;
    SLMap::iterator it = sl(arr)->upper_bound(std::string((const char*)key));
    if (it == sl(arr)->begin()) return NULL;
    return found(--it, key);
}
PPvoid_t JudySLPrev(Pcvoid_t arr, uint8_t* key, void*) {  // last key < *key
    if (!arr) return NULL;
    SLMap::iterator it = sl(arr)->lower_bound(std::string((const char*)key));
    if (it == sl(arr)->begin()) return NULL;
    return found(--it, key);
}
int JudySLDel(PPvoid_t arr, const uint8_t* key, void*) {
    if (!*arr) return 0;
    return (int)sl(*arr)->erase(std::string((const char*)key));
}
Word_t JudySLFreeArray(PPvoid_t arr, void*) {
    if (!*arr) return 0;
    Word_t n = 0;
    for (SLMap::iterator it = sl(*arr)->begin(); it != sl(*arr)->end(); ++it) n += it->first.size() + 1 + sizeof(Word_t);
    delete sl(*arr);
    *arr = NULL;
    return n;
}

PPvoid_t JudyLIns(PPvoid_t arr, Word_t idx, void*) {
    if (*arr == NULL) *arr = new LMap();
    return (PPvoid_t) & (*lm(*arr))[idx];
}
PPvoid_t JudyLGet(Pcvoid_t arr, Word_t idx, void*) {
    if (!arr) return NULL;
    LMap::iterator it = lm(arr)->find(idx);
    return it == lm(arr)->end() ? NULL : (PPvoid_t)&it->second;
}
PPvoid_t JudyLFirst(Pcvoid_t arr, Word_t* idx, void*) {
    if (!arr) return NULL;
    LMap::iterator it = lm(arr)->lower_bound(*idx);
    if (it == lm(arr)->end()) return NULL;
    *idx = it->first;
    return (PPvoid_t)&it->second;
}
PPvoid_t JudyLNext(Pcvoid_t arr, Word_t* idx, void*) {
    if (!arr) return NULL;
    LMap::iterator it = lm(arr)->upper_bound(*idx);
    if (it == lm(arr)->end()) return NULL;
    *idx = it->first;
    return (PPvoid_t)&it->second;
}
PPvoid_t JudyLLast(Pcvoid_t arr, Word_t* idx, void*) {
    if (!arr) return NULL;
    LMap::iterator it = lm(arr)->upper_bound(*idx);
    if (it == lm(arr)->begin()) return NULL;
    --it;
    *idx = it->first;
    return (PPvoid_t)&it->second;
}
PPvoid_t JudyLPrev(Pcvoid_t arr, Word_t* idx, void*) {
    if (!arr) return NULL;
    LMap::iterator it = lm(arr)->lower_bound(*idx);
    if (it == lm(arr)->begin()) return NULL;
    --it;
    *idx = it->first;
    return (PPvoid_t)&it->second;
}
int JudyLDel(PPvoid_t arr, Word_t idx, void*) {
    if (!*arr) return 0;
    return (int)lm(*arr)->erase(idx);
}
Word_t JudyLFreeArray(PPvoid_t arr, void*) {
    if (!*arr) return 0;
    Word_t n = lm(*arr)->size() * 2 * sizeof(Word_t);
    delete lm(*arr);
    *arr = NULL;
    return n;
}
}
