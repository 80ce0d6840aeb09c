// Cost of a first-touch page fault on this host (plain, MADV_HUGEPAGE, MADV_POPULATE_WRITE): the BAM
// decoder writes gigabytes of fresh batch memory, so this bounds what buffer recycling can save.
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#include <sys/mman.h>
static double now(){struct timespec a;clock_gettime(CLOCK_MONOTONIC,&a);return a.tv_sec+1e-9*a.tv_nsec;}
int main(){ size_t n=1ull<<30; 
 for(int mode=0;mode<3;++mode){ char*p=mmap(0,n,PROT_READ|PROT_WRITE,MAP_PRIVATE|MAP_ANONYMOUS,-1,0); if(mode==1) madvise(p,n,MADV_HUGEPAGE);
 double t=now(); if(mode==2){ madvise(p,n,23 /*MADV_POPULATE_WRITE*/);} for(size_t i=0;i<n;i+=4096)p[i]=1; double dt=now()-t; printf("mode %d: %.3f s, %.2f us/page\n",mode,dt,dt/(n/4096)*1e6); 
 t=now(); memset(p,1,n); printf("  memset warm %.3f s\n",now()-t); munmap(p,n);} }
