/* TEST INFRASTRUCTURE ONLY (oracle build).  Minimal stand-in for <windows.h>.
 *
 * The reference's evaluator.cpp:2 and marching.cpp:3 include <windows.h> only for
 * the Win32 file-dialog code (evaluator.cpp:240-332, marching.cpp:665-854).  None of
 * that is on the hot path; this header lets those two translation units compile
 * unmodified under g++ so that oracle/_ref can execute the real reference code.
 * Every dialog call reports "cancelled", so the file I/O bodies never run.
 */
#pragma once
#include <cstdio>
#include <cstdarg>
#include <cstddef>

typedef char TCHAR;
typedef int errno_t;
typedef unsigned long DWORD;
typedef unsigned short WORD;
typedef long LPARAM;
typedef void* HWND;
typedef void* HINSTANCE;
typedef const char* LPCSTR;
typedef char* LPSTR;

#ifndef MAX_PATH
#define MAX_PATH 260
#endif
#define TEXT(s) s
#define OFN_EXPLORER 0x00080000

struct OPENFILENAME {
    DWORD lStructSize;
    HWND hwndOwner;
    HINSTANCE hInstance;
    LPCSTR lpstrFilter;
    LPSTR lpstrCustomFilter;
    DWORD nMaxCustFilter;
    DWORD nFilterIndex;
    LPSTR lpstrFile;
    DWORD nMaxFile;
    LPSTR lpstrFileTitle;
    DWORD nMaxFileTitle;
    LPCSTR lpstrInitialDir;
    LPCSTR lpstrTitle;
    DWORD Flags;
    WORD nFileOffset;
    WORD nFileExtension;
    LPCSTR lpstrDefExt;
    LPARAM lCustData;
    void* lpfnHook;
    LPCSTR lpTemplateName;
};

static inline bool GetOpenFileName(OPENFILENAME*) { return false; }
static inline bool GetSaveFileName(OPENFILENAME*) { return false; }

static inline errno_t fopen_s(FILE** fp, const char* name, const char* mode) {
    *fp = std::fopen(name, mode);
    return *fp ? 0 : 1;
}
template <class T, size_t N> constexpr size_t mcb_shim_countof(T (&)[N]) { return N; }
#define _countof(a) mcb_shim_countof(a)
#define fscanf_s fscanf
#define fprintf_s fprintf
template <size_t N, class... A>
static inline int sprintf_s(char (&buf)[N], const char* fmt, A... a) {
    return std::snprintf(buf, N, fmt, a...);
}
