/* oracle/shim/windows.h -- TEST INFRASTRUCTURE ONLY.
 *
 * A minimal stand-in for <windows.h> so that the UNMODIFIED reference sources under
 * /root/reference (vadc.c, memory.h, string8.c, cembed.c) compile with gcc on Linux.
 * Only the handful of Win32 calls the reference's hot path and CLI touch are provided:
 *   memory.h:113      VirtualAlloc            -> calloc
 *   vadc.c:492-529    ReadFile/GetStdHandle   -> read(2) on fd 0
 *   vadc.c:835-842    QueryPerformance*       -> clock_gettime
 *   vadc.c:531-626    CreatePipe/CreateProcessW (ffmpeg child) -> stubbed, always fails
 *   string8.c:78-150  MultiByteToWideChar/WideCharToMultiByte -> byte-wise (ASCII) copies
 *   string8.c:193     CommandLineToArgvW/GetCommandLineW      -> /proc/self/cmdline
 * Nothing here is product code; it is used by oracle/Makefile to build oracle/_ref/.
 */
#ifndef VADC_ORACLE_SHIM_WINDOWS_H
#define VADC_ORACLE_SHIM_WINDOWS_H

#include <stddef.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <stdio.h>
#include <wchar.h>
#include <time.h>
#include <unistd.h>

typedef void *HANDLE;
typedef unsigned long DWORD; /* only used as a counter; width is irrelevant here */
typedef int BOOL;
typedef void *LPVOID;
typedef wchar_t *LPWSTR;
typedef const wchar_t *LPCWSTR;
typedef unsigned int UINT;

#ifndef TRUE
#define TRUE 1
#endif
#ifndef FALSE
#define FALSE 0
#endif

#define MEM_RESERVE 0x2000
#define MEM_COMMIT 0x1000
#define PAGE_READWRITE 0x04
#define CP_UTF8 65001
#define INVALID_HANDLE_VALUE ((HANDLE)(intptr_t)-1)
#define STD_INPUT_HANDLE ((DWORD)-10)
#define STD_OUTPUT_HANDLE ((DWORD)-11)
#define STD_ERROR_HANDLE ((DWORD)-12)
#define HANDLE_FLAG_INHERIT 1
#define STARTF_USESTDHANDLES 0x100

typedef union _LARGE_INTEGER
{
   long long QuadPart;
} LARGE_INTEGER;

typedef struct _SECURITY_ATTRIBUTES
{
   DWORD nLength;
   void *lpSecurityDescriptor;
   BOOL bInheritHandle;
} SECURITY_ATTRIBUTES;

typedef struct _STARTUPINFOW
{
   DWORD cb;
   DWORD dwFlags;
   HANDLE hStdInput;
   HANDLE hStdOutput;
   HANDLE hStdError;
} STARTUPINFOW, STARTUPINFO;

typedef struct _PROCESS_INFORMATION
{
   HANDLE hProcess;
   HANDLE hThread;
} PROCESS_INFORMATION;

static inline void *VirtualAlloc( void *address, size_t size, DWORD type, DWORD protect )
{
   (void)address; (void)type; (void)protect;
   return calloc( 1, size );
}

static inline HANDLE GetStdHandle( DWORD which )
{
   if ( which == STD_INPUT_HANDLE ) return (HANDLE)(intptr_t)(0 + 1);
   if ( which == STD_OUTPUT_HANDLE ) return (HANDLE)(intptr_t)(1 + 1);
   return (HANDLE)(intptr_t)(2 + 1);
}

/* Win32 ReadFile on a pipe: TRUE with *read>0 while data flows; FALSE (broken pipe) at EOF. */
static inline BOOL ReadFile( HANDLE h, void *buffer, DWORD to_read, DWORD *bytes_read, void *overlapped )
{
   (void)overlapped;
   int fd = (int)(intptr_t)h - 1;
   ssize_t n = read( fd, buffer, (size_t)to_read );
   if ( n <= 0 )
   {
      *bytes_read = 0;
      return FALSE;
   }
   *bytes_read = (DWORD)n;
   return TRUE;
}

static inline BOOL QueryPerformanceFrequency( LARGE_INTEGER *f )
{
   f->QuadPart = 1000000000LL;
   return TRUE;
}

static inline BOOL QueryPerformanceCounter( LARGE_INTEGER *c )
{
   struct timespec ts;
   clock_gettime( CLOCK_MONOTONIC, &ts );
   c->QuadPart = (long long)ts.tv_sec * 1000000000LL + ts.tv_nsec;
   return TRUE;
}

/* ffmpeg child process path (vadc.c:531-626) is not supported by the oracle build */
static inline BOOL CreatePipe( HANDLE *r, HANDLE *w, SECURITY_ATTRIBUTES *sa, DWORD size )
{
   (void)sa; (void)size;
   *r = INVALID_HANDLE_VALUE;
   *w = INVALID_HANDLE_VALUE;
   return FALSE;
}
static inline BOOL SetHandleInformation( HANDLE h, DWORD mask, DWORD flags ) { (void)h; (void)mask; (void)flags; return TRUE; }
static inline BOOL CloseHandle( HANDLE h ) { (void)h; return TRUE; }
static inline BOOL CreateProcessW( const wchar_t *app, wchar_t *cmd, void *pa, void *ta, BOOL inherit, DWORD flags,
                                   void *env, const wchar_t *cwd, STARTUPINFOW *si, PROCESS_INFORMATION *pi )
{
   (void)app; (void)cmd; (void)pa; (void)ta; (void)inherit; (void)flags; (void)env; (void)cwd; (void)si; (void)pi;
   return FALSE;
}

/* byte-wise conversions: sufficient for ASCII option names and file names */
static inline int MultiByteToWideChar( UINT cp, DWORD flags, const char *src, int src_len, wchar_t *dst, int dst_len )
{
   (void)cp; (void)flags;
   if ( src_len < 0 ) src_len = (int)strlen( src ) + 1;
   if ( dst == 0 || dst_len == 0 ) return src_len;
   int n = src_len < dst_len ? src_len : dst_len;
   for ( int i = 0; i < n; ++i ) dst[i] = (wchar_t)(unsigned char)src[i];
   return n;
}

static inline int WideCharToMultiByte( UINT cp, DWORD flags, const wchar_t *src, int src_len, char *dst, int dst_len,
                                       const char *def, BOOL *used_def )
{
   (void)cp; (void)flags; (void)def; (void)used_def;
   if ( src_len < 0 ) src_len = (int)wcslen( src ) + 1;
   if ( dst == 0 || dst_len == 0 ) return src_len;
   int n = src_len < dst_len ? src_len : dst_len;
   for ( int i = 0; i < n; ++i ) dst[i] = (char)src[i];
   return n;
}

static inline wchar_t *GetCommandLineW( void ) { return 0; }

/* argv from /proc/self/cmdline, widened byte-wise */
static inline wchar_t **CommandLineToArgvW( const wchar_t *cmdline, int *argc_out )
{
   (void)cmdline;
   static wchar_t *argv[256];
   static wchar_t storage[1 << 16];
   static char raw[1 << 16];
   int argc = 0;
   FILE *f = fopen( "/proc/self/cmdline", "rb" );
   size_t n = 0;
   if ( f )
   {
      n = fread( raw, 1, sizeof( raw ) - 1, f );
      fclose( f );
   }
   size_t pos = 0;
   size_t out = 0;
   while ( pos < n && argc < 255 )
   {
      size_t len = strlen( raw + pos );
      argv[argc++] = storage + out;
      for ( size_t i = 0; i <= len; ++i ) storage[out++] = (wchar_t)(unsigned char)raw[pos + i];
      pos += len + 1;
   }
   argv[argc] = 0;
   *argc_out = argc;
   return argv;
}

static inline void __debugbreak( void ) { __builtin_trap(); }

#endif /* VADC_ORACLE_SHIM_WINDOWS_H */
