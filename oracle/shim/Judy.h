/* Minimal Judy-compatible shim (test infrastructure, NOT product code).
 *
 * The reference links libJudy, which is not installed in this image. Judy is used only
 * on the BUILD path (annotation compression, src/annotation.c:918-1760,
 * src/replaceAnnotation.c:185-917, src/file_io.c:56-66); the query path never touches
 * it. This header provides the macro surface of <Judy.h> that those call sites use,
 * backed by ordered maps in judy_shim.cpp:
 *   JudySL = ordered map  NUL-terminated byte string (unsigned byte order) -> Word_t
 *   JudyL  = ordered map  Word_t -> Word_t
 * Value slots are zero-initialised on insert and their addresses stay valid until
 * deletion. Macros are brace blocks because the reference writes
 * `if (c) JSLD(...) else {` (src/annotation.c:1517).
 */
#pragma once
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef unsigned long Word_t;
typedef Word_t* PWord_t;
typedef void* Pvoid_t;
typedef void** PPvoid_t;
typedef const void* Pcvoid_t;

#define PJERR ((Pvoid_t)(~0UL))
#define PPJERR ((PPvoid_t)(~0UL))
#define PJE0 ((void*)0)
#define JERR (-1)

PPvoid_t JudySLIns(PPvoid_t arr, const uint8_t* key, void* err);
PPvoid_t JudySLGet(Pcvoid_t arr, const uint8_t* key, void* err);
PPvoid_t JudySLFirst(Pcvoid_t arr, uint8_t* key, void* err);
PPvoid_t JudySLNext(Pcvoid_t arr, uint8_t* key, void* err);
PPvoid_t JudySLLast(Pcvoid_t arr, uint8_t* key, void* err);
PPvoid_t JudySLPrev(Pcvoid_t arr, uint8_t* key, void* err);
int JudySLDel(PPvoid_t arr, const uint8_t* key, void* err);
Word_t JudySLFreeArray(PPvoid_t arr, void* err);

PPvoid_t JudyLIns(PPvoid_t arr, Word_t idx, void* err);
PPvoid_t JudyLGet(Pcvoid_t arr, Word_t idx, void* err);
PPvoid_t JudyLFirst(Pcvoid_t arr, Word_t* idx, void* err);
PPvoid_t JudyLNext(Pcvoid_t arr, Word_t* idx, void* err);
PPvoid_t JudyLLast(Pcvoid_t arr, Word_t* idx, void* err);
PPvoid_t JudyLPrev(Pcvoid_t arr, Word_t* idx, void* err);
int JudyLDel(PPvoid_t arr, Word_t idx, void* err);
Word_t JudyLFreeArray(PPvoid_t arr, void* err);

#ifdef __cplusplus
}
#endif

#define JSLI(PV, PArray, Index) { (PV) = (__typeof__(PV))JudySLIns((PPvoid_t)&(PArray), (const uint8_t*)(Index), PJE0); }
#define JSLG(PV, PArray, Index) { (PV) = (__typeof__(PV))JudySLGet((Pcvoid_t)(PArray), (const uint8_t*)(Index), PJE0); }
#define JSLF(PV, PArray, Index) { (PV) = (__typeof__(PV))JudySLFirst((Pcvoid_t)(PArray), (uint8_t*)(Index), PJE0); }
#define JSLN(PV, PArray, Index) { (PV) = (__typeof__(PV))JudySLNext((Pcvoid_t)(PArray), (uint8_t*)(Index), PJE0); }
#define JSLL(PV, PArray, Index) { (PV) = (__typeof__(PV))JudySLLast((Pcvoid_t)(PArray), (uint8_t*)(Index), PJE0); }
#define JSLP(PV, PArray, Index) { (PV) = (__typeof__(PV))JudySLPrev((Pcvoid_t)(PArray), (uint8_t*)(Index), PJE0); }
#define JSLD(Rc, PArray, Index) { (Rc) = (__typeof__(Rc))(intptr_t)JudySLDel((PPvoid_t)&(PArray), (const uint8_t*)(Index), PJE0); }
#define JSLFA(Rc, PArray)       { (Rc) = (__typeof__(Rc))JudySLFreeArray((PPvoid_t)&(PArray), PJE0); }

#define JLI(PV, PArray, Index)  { (PV) = (__typeof__(PV))JudyLIns((PPvoid_t)&(PArray), (Word_t)(Index), PJE0); }
#define JLG(PV, PArray, Index)  { (PV) = (__typeof__(PV))JudyLGet((Pcvoid_t)(PArray), (Word_t)(Index), PJE0); }
#define JLF(PV, PArray, Index)  { (PV) = (__typeof__(PV))JudyLFirst((Pcvoid_t)(PArray), (Word_t*)&(Index), PJE0); }
#define JLN(PV, PArray, Index)  { (PV) = (__typeof__(PV))JudyLNext((Pcvoid_t)(PArray), (Word_t*)&(Index), PJE0); }
#define JLL(PV, PArray, Index)  { (PV) = (__typeof__(PV))JudyLLast((Pcvoid_t)(PArray), (Word_t*)&(Index), PJE0); }
#define JLP(PV, PArray, Index)  { (PV) = (__typeof__(PV))JudyLPrev((Pcvoid_t)(PArray), (Word_t*)&(Index), PJE0); }
#define JLD(Rc, PArray, Index)  { (Rc) = (__typeof__(Rc))(intptr_t)JudyLDel((PPvoid_t)&(PArray), (Word_t)(Index), PJE0); }
#define JLFA(Rc, PArray)        { (Rc) = (__typeof__(Rc))JudyLFreeArray((PPvoid_t)&(PArray), PJE0); }
