// Throw-away experiment (not product, not test): upper bound on what SURVEY 8a row a12 (block splitting,
// deflate_compress.c:2091-2218) could save here.  Parses the emulator's own members back into tokens and compares the
// entropy cost of one DEFLATE block with the best of 7 two-block splits (60-byte header per block assumed).
//   g++ -O2 -o build/split_gain tools/experiments/split_gain.cpp && build/split_gain 6 <file>
// Measured on the 8 MiB corpora at level 6: FASTQ-like 0.25 %, SAM-like 0.00 %.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <cmath>
#include <algorithm>
#include <numeric>
#define main emul_main_unused
#include "../../tests/model/emul.cpp"
#undef main
struct Tok { uint32_t pos; uint16_t lsym; int16_t dsym; uint8_t xbits; };
static double cost(const std::vector<Tok>&t, size_t a, size_t b){ // entropy bits of tokens [a,b) + extra bits
  double f[288]={0}, d[32]={0}; double x=0; for(size_t i=a;i<b;i++){ f[t[i].lsym]++; if(t[i].dsym>=0) d[t[i].dsym]++; x+=t[i].xbits; } f[256]+=1;
  double tot=0, td=0; for(double v:f) tot+=v; for(double v:d) td+=v; double bits=x; for(double v:f) if(v>0) bits+=v*std::log2(tot/v); for(double v:d) if(v>0) bits+=v*std::log2(td/v); return bits; }
int main(int argc,char**argv){
  int level=atoi(argv[1]); FILE*f=fopen(argv[2],"rb"); std::vector<uint8_t> in; {uint8_t buf[65536]; size_t r; while((r=fread(buf,1,sizeof buf,f))>0) in.insert(in.end(),buf,buf+r);} fclose(f);
  double one=0, two=0, real=0; int nsplit=0, nb=0; const double HDR=60*8;
  for(size_t off=0; off+65280<=in.size() && nb<40; off+=65280, nb++){
    uint8_t dst[65536]; uint32_t dl=0; bgemul_compress_block(in.data()+off,65280,level,0,dst,&dl); real+=dl;
    // decode tokens with a tiny inflate of dst (one dynamic block expected) using zlib? simpler: reuse emulator state: static Emu not accessible; re-derive tokens by parsing the deflate stream
    // --- minimal inflate that records tokens ---
    const uint8_t*p=dst+18; uint64_t bitpos=0; auto get=[&](int n){ uint32_t v=0; for(int i=0;i<n;i++){ v|=((p[bitpos>>3]>>(bitpos&7))&1u)<<i; bitpos++; } return v; };
    std::vector<Tok> toks; uint32_t pos=0; bool last=false;
    static const uint16_t lbase[29]={3,4,5,6,7,8,9,10,11,13,15,17,19,23,27,31,35,43,51,59,67,83,99,115,131,163,195,227,258}; static const uint8_t lex[29]={0,0,0,0,0,0,0,0,1,1,1,1,2,2,2,2,3,3,3,3,4,4,4,4,5,5,5,5,0};
    static const uint8_t dex[30]={0,0,0,0,1,1,2,2,3,3,4,4,5,5,6,6,7,7,8,8,9,9,10,10,11,11,12,12,13,13};
    while(!last){ last=get(1); int bt=get(2); if(bt!=2){ toks.clear(); break; }
      int nl=get(5)+257, nd=get(5)+1, np=get(4)+4; static const int ord[19]={16,17,18,0,8,7,9,6,10,5,11,4,12,3,13,2,14,1,15}; uint8_t pl[19]={0}; for(int i=0;i<np;i++) pl[ord[i]]=get(3);
      auto build=[&](const uint8_t*lens,int n,std::vector<int>&codes,std::vector<int>&syms){ // canonical decode table as list
        int cnt[16]={0}; for(int i=0;i<n;i++) cnt[lens[i]]++; cnt[0]=0; int next[16]; int code=0; for(int l=1;l<16;l++){ code=(code+cnt[l-1])<<1; next[l]=code; } codes.assign(n,0); for(int i=0;i<n;i++) if(lens[i]) codes[i]=next[lens[i]]++; (void)syms; };
      auto dec=[&](const uint8_t*lens,int n,const std::vector<int>&codes){ int code=0; for(int l=1;l<16;l++){ code=(code<<1)|get(1); for(int s=0;s<n;s++) if(lens[s]==l&&codes[s]==code) return s; } return -1; };
      std::vector<int> pc,dummy; build(pl,19,pc,dummy); uint8_t ll[320]={0}; int i=0; while(i<nl+nd){ int s=dec(pl,19,pc); if(s<16) ll[i++]=s; else if(s==16){ int r=3+get(2); while(r--) { ll[i]=ll[i-1]; i++; } } else if(s==17){ i+=3+get(3);} else { i+=11+get(7);} }
      std::vector<int> lc,dc; build(ll,nl,lc,dummy); build(ll+nl,nd,dc,dummy);
      for(;;){ int s=dec(ll,nl,lc); if(s==256) break; Tok t; t.pos=pos; t.lsym=s; t.dsym=-1; t.xbits=0; if(s<256){ pos++; } else { int ls=s-257; int len=lbase[ls]+get(lex[ls]); int ds=dec(ll+nl,nd,dc); get(dex[ds]); t.dsym=ds; t.xbits=lex[ls]+dex[ds]; pos+=len; } toks.push_back(t); }
    }
    if(toks.empty()) continue;
    double c1=cost(toks,0,toks.size())+HDR; double best=c1; 
    for(int k=1;k<8;k++){ uint32_t sp=65280*k/8; size_t i=std::lower_bound(toks.begin(),toks.end(),sp,[](const Tok&a,uint32_t v){return a.pos<v;})-toks.begin(); double c2=cost(toks,0,i)+cost(toks,i,toks.size())+2*HDR; best=std::min(best,c2); }
    one+=c1; two+=best; if(best<c1) nsplit++;
  }
  printf("blocks %d: entropy-cost one block %.0f bytes, best of 7 split points %.0f bytes (%.2f%% smaller), %d blocks would split; real size %.0f\n", nb, one/8, two/8, 100*(one-two)/one, nsplit, real);
}
