// Round-count simulator of the lock-step sweep (host C, exact claims, no hash collisions): batches of 512 bonds,
// earliest-bond-wins claim rounds, star merging into the K largest clusters, tail mode.
//   gcc -O2 -o sim scripts/sim_rounds.c;  ./sim L policy K tailmax [seed]      (policy 0: one hub; 1: K hubs, THR=<min size> env)
// Reproduces 825 rounds / 213 tails per L=256 run (831 / 215 measured in the kernel).  See profiles/sweep_cta_shape_r1.txt.
// round-count simulator of the lock-step sweep (exact claims, no hash collisions)
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <stdint.h>
#define T 512
static int *par, *sz;
static int find(int x){ while(par[x]!=x){ par[x]=par[par[x]]; x=par[x]; } return x; }
static uint64_t rng=88172645463325252ull;
static uint64_t xr(){ rng^=rng<<13; rng^=rng>>7; rng^=rng<<17; return rng; }
int main(int argc,char**argv){
  int L=atoi(argv[1]); int policy=atoi(argv[2]); int K=argc>3?atoi(argv[3]):1; int tailmax=argc>4?atoi(argv[4]):32;
  rng ^= (argc>5? strtoull(argv[5],0,10)*0x9E3779B97F4A7C15ull:0);
  int N=L*L, M=2*L*(L-1);
  int *eu=malloc(4*M),*ev=malloc(4*M); int m=0;
  for(int x=0;x<L;x++)for(int y=0;y<L;y++){ int id=x*L+y; if(y+1<L){eu[m]=id;ev[m]=id+1;m++;} if(x+1<L){eu[m]=id;ev[m]=id+L;m++;} }
  int *perm=malloc(4*M); for(int i=0;i<M;i++)perm[i]=i; for(int i=M-1;i>0;i--){int j=xr()%(i+1);int t=perm[i];perm[i]=perm[j];perm[j]=t;}
  par=malloc(4*N); sz=malloc(4*N); for(int i=0;i<N;i++){par[i]=i;sz[i]=1;}
  int *owner=malloc(4*N); for(int i=0;i<N;i++)owner[i]=1<<30;
  long rounds_cta=0, rounds_tail=0, tails=0, batches=0; long hist[64]={0};
  int hubs[8]; int nh=0;
  long rounds_by_decile[10]={0};
  for(int n0=0;n0<M;n0+=T){
    int cnt = M-n0<T?M-n0:T; int ru[T],rv[T],pend[T];
    for(int i=0;i<cnt;i++){ int e=perm[n0+i]; ru[i]=find(eu[e]); rv[i]=find(ev[e]); pend[i]=ru[i]!=rv[i]; }
    int r=0; int intail=0;
    for(;;){
      int left=0; for(int i=0;i<cnt;i++){ if(pend[i]){ ru[i]=find(ru[i]); rv[i]=find(rv[i]); pend[i]=ru[i]!=rv[i]; } left+=pend[i]; }
      if(!left)break;
      if(left<=tailmax && !intail){ intail=1; tails++; }
      // hubs: K largest roots among ... approximate: track by scanning touched roots + previous hubs
      // exact: find the K largest clusters overall (cheap enough: maintain via scan every round of candidates)
      // candidates = previous hubs (re-found) + roots touched in this batch
      int cand[2*T+8]; int nc=0; for(int h=0;h<nh;h++)cand[nc++]=find(hubs[h]); for(int i=0;i<cnt;i++){cand[nc++]=find(ru[i]);cand[nc++]=find(rv[i]);}
      nh=0; for(int k=0;k<K;k++){ int best=-1; for(int c=0;c<nc;c++){ int x=cand[c]; int dup=0; for(int h=0;h<nh;h++) if(hubs[h]==x)dup=1; if(dup)continue; if(best<0||sz[x]>sz[best]||(sz[x]==sz[best]&&x>best))best=x; } if(best>=0)hubs[nh++]=best; }
      if(policy==0){ nh = nh>1?1:nh; } else { int thr=atoi(getenv("THR")?getenv("THR"):"0"); int k2=1; for(int h=1;h<nh;h++) if(sz[hubs[h]]>=thr) hubs[k2++]=hubs[h]; nh=k2; }
      // claims
      int star[T], o[T], hb[T];
      for(int i=0;i<cnt;i++){ star[i]=0; if(!pend[i])continue; int hu=-1,hv=-1; for(int h=0;h<nh;h++){ if(ru[i]==hubs[h])hu=h; if(rv[i]==hubs[h])hv=h; }
        if(hu>=0 && hv>=0){ /* hub-hub bond: normal bond claiming both */ star[i]=0; }
        else if(hu>=0){ star[i]=1; o[i]=rv[i]; hb[i]=hu; } else if(hv>=0){ star[i]=1; o[i]=ru[i]; hb[i]=hv; } }
      for(int i=0;i<cnt;i++){ if(!pend[i])continue; if(star[i]){ if(owner[o[i]]>i)owner[o[i]]=i; } else { if(owner[ru[i]]>i)owner[ru[i]]=i; if(owner[rv[i]]>i)owner[rv[i]]=i; } }
      int own[T]; int bmin=1<<30;
      // a hub that is claimed by a normal (hub-hub) bond j: star bonds of that hub with index > j are blocked; and the hub-hub bond must wait for earlier star bonds of its hubs (they merge this round only if < bmin) -> treat: hub-hub bond owns iff it owns both claims AND no earlier pending star bond on either hub
      int firststar[8]; for(int h=0;h<8;h++)firststar[h]=1<<30;
      for(int i=0;i<cnt;i++) if(pend[i]&&star[i]&&firststar[hb[i]]>i) firststar[hb[i]]=i;
      for(int i=0;i<cnt;i++){ own[i]=0; if(!pend[i])continue; if(star[i]){ own[i]= owner[o[i]]==i && owner[hubs[hb[i]]]>i; } else { own[i]= owner[ru[i]]==i && owner[rv[i]]==i; if(own[i]) for(int h=0;h<nh;h++) if((ru[i]==hubs[h]||rv[i]==hubs[h]) && firststar[h]<i) own[i]=0; }
        if(!own[i] && bmin>i) bmin=i; }
      // merges: non-star owners
      int merged=0;
      for(int i=0;i<cnt;i++){ if(!pend[i]||!own[i]||star[i])continue; int a=ru[i],b=rv[i]; if(sz[a]<sz[b]){int t=a;a=b;b=t;} par[b]=a; sz[a]+=sz[b]; pend[i]=0; merged++; }
      for(int i=0;i<cnt;i++){ if(!pend[i]||!own[i]||!star[i])continue; if(i>=bmin)continue; int h=hubs[hb[i]]; /* h stays root? require sz[h]>=sz[o] else becomes root swap; fine for counting */ int a=find(h),b=o[i]; if(a==b){pend[i]=0;continue;} if(sz[a]<sz[b]){int t=a;a=b;b=t;} par[b]=a; sz[a]+=sz[b]; pend[i]=0; merged++; }
      for(int i=0;i<cnt;i++){ if(ru[i]>=0){ owner[ru[i]]=1<<30; owner[rv[i]]=1<<30; } }
      for(int h=0;h<nh;h++) owner[hubs[h]]=1<<30;
      if(!merged){ fprintf(stderr,"stuck batch %d round %d left %d\n",n0/T,r,left); return 1; }
      r++; if(intail)rounds_tail++; else rounds_cta++;
    }
    batches++; hist[r>63?63:r]++; rounds_by_decile[(int)(10.0*n0/M)]+=r;
  }
  printf("L=%d policy=%d K=%d tailmax=%d: batches %ld, CTA rounds %ld, tail rounds %ld (tails %ld), total %ld\n",L,policy,K,tailmax,batches,rounds_cta,rounds_tail,tails,rounds_cta+rounds_tail);
  printf("rounds by decile of n/M:"); for(int d=0;d<10;d++)printf(" %ld",rounds_by_decile[d]); printf("\n");
  return 0;
}
