// Mutation fuzz of kdf::inflate_raw against zlib (build with -fsanitize=address,undefined):
//   g++ -std=c++17 -O1 -g -fsanitize=address,undefined -o /tmp/fuzz_infl scripts/fuzz_inflate.cpp \
//       kmer_denovo_filter_b200/csrc/kdf_inflate.cpp -lz && /tmp/fuzz_infl some.bam [seed] [trials]
// Every mutated stream must be rejected by both decoders or inflate to the same bytes.
#include <zlib.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <stdint.h>
#include <vector>
#include <random>
#include "/root/repo/kmer_denovo_filter_b200/csrc/kdf_inflate.h"
static int zinf(const uint8_t* in, size_t n, uint8_t* out, size_t on){ z_stream zs; memset(&zs,0,sizeof zs); inflateInit2(&zs,-15); zs.next_in=(Bytef*)in; zs.avail_in=n; zs.next_out=out; zs.avail_out=on; int rc=inflate(&zs,Z_FINISH); size_t to=zs.total_out, ti=zs.total_in; inflateEnd(&zs); return (rc==Z_STREAM_END && to==on) ? (ti==n?1:2) : 0; }
int main(int argc,char**argv){ FILE*f=fopen(argv[1],"rb");fseek(f,0,SEEK_END);long n=ftell(f);fseek(f,0,SEEK_SET);std::vector<uint8_t> b(n);if(fread(b.data(),1,n,f)!=(size_t)n)return 1;
 std::mt19937_64 rng(argc>2?atoi(argv[2]):1); long trials=argc>3?atol(argv[3]):20000;
 std::vector<std::pair<long,int>> blocks; long p=0; while(p<n){int bs=(b[p+16]|(b[p+17]<<8))+1; blocks.push_back({p,bs}); p+=bs;}
 long acc_both=0, rej_both=0, only_z=0, only_k=0, diff=0, trailing=0;
 for(long t=0;t<trials;++t){ auto bl=blocks[rng()%blocks.size()]; size_t clen=bl.second-26; uint32_t isize; memcpy(&isize,&b[bl.first+bl.second-4],4);
   size_t ilen = clen; int kind=rng()%6;
   if(kind==4 && clen>4) ilen = rng()%clen;          // truncated input
   uint8_t* in=(uint8_t*)malloc(ilen?ilen:1); memcpy(in,&b[bl.first+18],ilen);
   if(kind<=2 && ilen){ int nm=1+rng()%3; for(int k=0;k<nm;++k){ size_t at = (kind==0)? rng()%std::min<size_t>(ilen,40) : rng()%ilen; in[at]^= (uint8_t)(1u<<(rng()%8)); } }
   if(kind==3 && ilen){ size_t at=rng()%ilen; in[at]=(uint8_t)rng(); }
   size_t on=isize; if(kind==5) on = (rng()%2)? isize+1+rng()%100 : (isize>0? rng()%isize:0);   // wrong expected size
   uint8_t* o1=(uint8_t*)malloc(on?on:1); uint8_t* o2=(uint8_t*)malloc(on?on:1);
   bool k=kdf::inflate_raw(in,ilen,o1,on); int z=zinf(in,ilen,o2,on);
   if(k&&z){ acc_both++; if(memcmp(o1,o2,on)) diff++; if(z==2) trailing++; } else if(!k&&!z) rej_both++; else if(z) only_z++; else only_k++;
   free(in);free(o1);free(o2);} 
 printf("trials %ld: both accept %ld (different output %ld, zlib left input unread %ld), both reject %ld, only zlib %ld, only kdf %ld\n",trials,acc_both,diff,trailing,rej_both,only_z,only_k); return diff?1:0;}
