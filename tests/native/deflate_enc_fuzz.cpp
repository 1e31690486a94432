// Probe (not part of the pytest suites): the PNG encoder's DEFLATE writer (deflate_enc_core.h, host build) on random inputs of five
// kinds, every level, against zlib's uncompress, under ASan + UBSan:
//   g++ -O1 -g -std=c++17 -fsanitize=address,undefined -fno-sanitize=shift-base tests/native/deflate_enc_fuzz.cpp -o /tmp/asan/defl_fuzz -lz && /tmp/asan/defl_fuzz 400
// (round 2: 400 streams of up to 140 KB, clean)
#include <cstdio>
#include <cstdlib>
#include <random>
#include <vector>
#include <zlib.h>
#include "deflate_enc_sim.cpp"
int main(int argc,char**argv){ long iters=argc>1?atol(argv[1]):300; std::mt19937 rng(3); long bad=0;
 for(long it=0;it<iters;it++){ size_t n=rng()%140000; std::vector<uint8_t> d(n); int kind=rng()%5;
  for(size_t i=0;i<n;i++) d[i]= kind==0?(uint8_t)rng(): kind==1?(uint8_t)(rng()%3): kind==2?(uint8_t)(i%7==0?rng():d[i?i-1:0]) : kind==3? (uint8_t)((int)(rng()%9)-4) : (uint8_t)(i*31>>3);
  int lvl=rng()%10; std::vector<uint8_t> out(n+n/500+1024); long m=defenc_compress(d.data(),(long)n,lvl,out.data(),(long)out.size()); if(m<=0){printf("compress failed n=%zu lvl=%d\n",n,lvl);bad++;continue;}
  std::vector<uint8_t> back(n+1); uLongf bl=(uLongf)back.size(); int rc=uncompress(back.data(),&bl,out.data(),(uLong)m); if(rc!=Z_OK||bl!=n||memcmp(back.data(),d.data(),n)){printf("roundtrip failed n=%zu lvl=%d rc=%d\n",n,lvl,rc);bad++;}
 }
 printf("%ld iterations, %ld bad\n",iters,bad); return bad!=0; }
