// Mutation fuzzer for libbmfr_io's EXR reader and camera-header parser (run under ASan + UBSan):
//   g++ -std=c++17 -g -O1 -fsanitize=address,undefined -fno-sanitize-recover=undefined scripts/fuzz_io.cpp \
//       bmfr_b200/csrc/bmfr_io.cpp -lz -o /tmp/fuzz_io
//   /tmp/fuzz_io SEED ITERATIONS tests/golden/exr/*.exr some_camera_matrices.h
// Scratch files go to /dev/shm.  r01: 1.25 M mutated inputs (all five compressions) clean after the data-window overflow fix.
#include "../include/bmfr_io.h"
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <vector>
#include <string>
static std::vector<unsigned char> slurp(const char* p){FILE*f=fopen(p,"rb");std::vector<unsigned char> v;if(!f)return v;fseek(f,0,SEEK_END);long n=ftell(f);fseek(f,0,SEEK_SET);v.resize(n);fread(v.data(),1,n,f);fclose(f);return v;}
int main(int argc,char**argv){
  unsigned seed=atoi(argv[1]); int iters=atoi(argv[2]); srand(seed);
  std::vector<std::vector<unsigned char>> seeds; for(int i=3;i<argc;++i) seeds.push_back(slurp(argv[i]));
  std::vector<float> out(64*64*3);
  int ok=0;
  for(int it=0;it<iters;++it){
    if(getenv("FZ_TRACE")) fprintf(stderr,"it %d\n",it);
    std::vector<unsigned char> b=seeds[rand()%seeds.size()];
    int muts=1+rand()%6;
    for(int m=0;m<muts;++m){
      int kind=rand()%5; size_t pos=b.empty()?0:rand()%b.size();
      if(kind==0&&!b.empty()) b[pos]^=1<<(rand()%8);
      else if(kind==1&&!b.empty()) b[pos]=rand();
      else if(kind==2&&!b.empty()) b.resize(pos);
      else if(kind==3&&b.size()>8){ int v=(rand()%3==0)?0x7fffffff:(rand()%3==0?-1:rand()); memcpy(&b[pos%(b.size()-4)],&v,4);}
      else if(kind==4&&!b.empty()) b.insert(b.begin()+pos,(unsigned char)rand());
    }
    FILE*f=fopen("/dev/shm/bmfr_fuzz_cur.exr","wb"); fwrite(b.data(),1,b.size(),f); fclose(f);
    int w=0,h=0,c=0;
    if(bmfr_io_exr_info("/dev/shm/bmfr_fuzz_cur.exr",&w,&h,&c)==0 && w>0&&h>0&&(long)w*h<=64*64){
      if(bmfr_io_read_exr_rgb("/dev/shm/bmfr_fuzz_cur.exr",w,h,out.data())==0) ++ok;
    } else { bmfr_io_read_exr_rgb("/dev/shm/bmfr_fuzz_cur.exr",24,19,out.data()); }
    // header parser on the same bytes as text
    float m[32],o[4],pl,nl; int nm,no;
    FILE*g=fopen("/dev/shm/bmfr_fuzz_cur.h","wb"); fwrite(b.data(),1,b.size(),g); fclose(g);
    bmfr_io_parse_camera_header("/dev/shm/bmfr_fuzz_cur.h",2,m,o,&nm,&no,&pl,&nl);
  }
  printf("iterations %d, still-valid files %d\n",iters,ok);
  return 0;}
