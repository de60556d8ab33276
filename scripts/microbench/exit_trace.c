// LD_PRELOAD shim: prints a backtrace when the process calls exit / _exit (debugging aid)
#define _GNU_SOURCE
#include <execinfo.h>
#include <stdio.h>
#include <sys/syscall.h>
#include <unistd.h>
static void bt(const char *who, int code)
{
    void *buf[64];
    int n = backtrace(buf, 64);
    dprintf(2, "== %s(%d) called from:\n", who, code);
    backtrace_symbols_fd(buf, n, 2);
}
void _exit(int code) { bt("_exit", code); syscall(SYS_exit_group, code); for (;;) {} }
void _Exit(int code) { bt("_Exit", code); syscall(SYS_exit_group, code); for (;;) {} }
#include <dlfcn.h>
void exit(int code)
{
    bt("exit", code);
    void (*real)(int) = (void (*)(int))dlsym(RTLD_NEXT, "exit");
    real(code);
    for (;;) {}
}
