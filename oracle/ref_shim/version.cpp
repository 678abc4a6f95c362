// toast_sys_environment.cpp expects this symbol from the generated version file.
extern "C" { const char* TOAST_VERSION = "oracle"; }
