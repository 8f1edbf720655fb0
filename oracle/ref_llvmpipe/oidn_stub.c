/* OIDN is binary-only/Windows-only in the reference; denoise is never invoked by the harness */
void* oidnNewDevice(int t){return 0;} void oidnCommitDevice(void* d){} void oidnReleaseDevice(void* d){}
int oidnGetDeviceError(void* d,const char** m){return 0;} void* oidnNewFilter(void* d,const char* t){return 0;} void oidnReleaseFilter(void* f){}
void oidnSetSharedFilterImage(void* f,const char* n,void* p,int fmt,unsigned long w,unsigned long h,unsigned long o,unsigned long ps,unsigned long rs){}
void oidnSetFilter1b(void* f,const char* n,int v){} void oidnCommitFilter(void* f){} void oidnExecuteFilter(void* f){}
