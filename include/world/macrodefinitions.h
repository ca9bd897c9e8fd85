/* world-b200 drop-in header.  Replaces externs/WORLD_v2/src/world/macrodefinitions.h:66-74
 * (C linkage macros) of the reference; same macro names so existing callers compile. */
#ifndef WORLD_MACRODEFINITIONS_H_
#define WORLD_MACRODEFINITIONS_H_
#undef WORLD_BEGIN_C_DECLS
#undef WORLD_END_C_DECLS
#ifdef __cplusplus
#define WORLD_BEGIN_C_DECLS extern "C" {
#define WORLD_END_C_DECLS }
#else
#define WORLD_BEGIN_C_DECLS
#define WORLD_END_C_DECLS
#endif
#if defined(__GNUC__) && __GNUC__ >= 4
#define WORLD_API __attribute__((visibility("default")))
#else
#define WORLD_API
#endif
#endif
