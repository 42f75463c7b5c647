// Variant of sim_rounds.c: the <= carrymax pending bonds of a batch are carried into the next batch as its earliest bonds.
//   gcc -O2 -o sim2 scripts/sim_rounds_carry.c;  ./sim2 L carrymax tailmax [seed]
// carry-over simulator: pending bonds (<= CARRY) of a batch are carried into the next batch as its earliest elements
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <stdint.h>
#define T 512
#define CMAX 64
static int *par, *sz;
static int find(int x){ while(par[x]!=x){ par[x]=par[par[x]]; x=par[x]; } return x; }
static uint64_t rng=88172645463325252ull;
static uint64_t xr(){ rng^=rng<<13; rng^=rng>>7; rng^=rng<<17; return rng; }
int main(int argc,char**argv){
  int L=atoi(argv[1]); int carrymax=atoi(argv[2]); int tailmax=atoi(argv[3]);
  rng ^= (argc>4? strtoull(argv[4],0,10)*0x9E3779B97F4A7C15ull:0);
  int N=L*L, M=2*L*(L-1);
  int *eu=malloc(4*M),*ev=malloc(4*M); int m=0;
  for(int x=0;x<L;x++)for(int y=0;y<L;y++){ int id=x*L+y; if(y+1<L){eu[m]=id;ev[m]=id+1;m++;} if(x+1<L){eu[m]=id;ev[m]=id+L;m++;} }
  int *perm=malloc(4*M); for(int i=0;i<M;i++)perm[i]=i; for(int i=M-1;i>0;i--){int j=xr()%(i+1);int t=perm[i];perm[i]=perm[j];perm[j]=t;}
  par=malloc(4*N); sz=malloc(4*N); for(int i=0;i<N;i++){par[i]=i;sz[i]=1;}
  int *owner=malloc(4*N); for(int i=0;i<N;i++)owner[i]=1<<30;
  long rounds_cta=0, rounds_tail=0, tails=0, batches=0, carried_tot=0, carries=0;
  int hub=0;
  int ru[T+CMAX],rv[T+CMAX],pend[T+CMAX]; int ncar=0;
  // verification of sequential semantics: record (size a,size b) per bond and compare with sequential run
  for(int n0=0;n0<M || ncar>0;n0+=T){
    int cnt = n0<M ? (M-n0<T?M-n0:T) : 0; int tot=ncar+cnt;
    for(int i=0;i<cnt;i++){ int e=perm[n0+i]; ru[ncar+i]=find(eu[e]); rv[ncar+i]=find(ev[e]); pend[ncar+i]=ru[ncar+i]!=rv[ncar+i]; }
    int last = n0+T>=M;
    int intail=0;
    for(;;){
      int left=0; for(int i=0;i<tot;i++){ if(pend[i]){ ru[i]=find(ru[i]); rv[i]=find(rv[i]); pend[i]=ru[i]!=rv[i]; } left+=pend[i]; }
      if(!left)break;
      if(!last && left<=carrymax) break;       // carry them
      if(left<=tailmax && !intail){ intail=1; tails++; }
      // hub = largest cluster
      for(int i=0;i<tot;i++){ if(sz[ru[i]]>sz[find(hub)])hub=ru[i]; if(sz[rv[i]]>sz[find(hub)])hub=rv[i]; } hub=find(hub);
      int star[T+CMAX], o[T+CMAX];
      for(int i=0;i<tot;i++){ star[i]=0; if(!pend[i])continue; if(ru[i]==hub){star[i]=1;o[i]=rv[i];} else if(rv[i]==hub){star[i]=1;o[i]=ru[i];} }
      for(int i=0;i<tot;i++){ if(!pend[i])continue; if(star[i]){ if(owner[o[i]]>i)owner[o[i]]=i; } else { if(owner[ru[i]]>i)owner[ru[i]]=i; if(owner[rv[i]]>i)owner[rv[i]]=i; } }
      int own[T+CMAX]; int bmin=1<<30;
      for(int i=0;i<tot;i++){ own[i]=0; if(!pend[i])continue; if(star[i]) own[i]= owner[o[i]]==i; else own[i]= owner[ru[i]]==i && owner[rv[i]]==i; if(!own[i] && bmin>i) bmin=i; }
      int merged=0;
      for(int i=0;i<tot;i++){ if(!pend[i]||!own[i]||star[i])continue; int a=ru[i],b=rv[i]; if(sz[a]<sz[b]){int t=a;a=b;b=t;} par[b]=a; sz[a]+=sz[b]; pend[i]=0; merged++; }
      for(int i=0;i<tot;i++){ if(!pend[i]||!own[i]||!star[i])continue; if(i>=bmin)continue; int a=find(hub),b=o[i]; if(sz[a]<sz[b]){int t=a;a=b;b=t;} par[b]=a; sz[a]+=sz[b]; pend[i]=0; merged++; }
      for(int i=0;i<tot;i++){ owner[ru[i]]=1<<30; owner[rv[i]]=1<<30; }
      if(!merged){ fprintf(stderr,"stuck\n"); return 1; }
      if(intail)rounds_tail++; else rounds_cta++;
    }
    // compact the pending bonds (in order) to the front
    int k=0; for(int i=0;i<tot;i++) if(pend[i]){ ru[k]=ru[i]; rv[k]=rv[i]; pend[k]=1; k++; }
    ncar=k; if(k){carries++; carried_tot+=k;}
    batches++;
    if(n0>=M && ncar==0)break;
  }
  printf("L=%d carrymax=%d tailmax=%d: batches %ld, CTA rounds %ld, tail rounds %ld (tails %ld), carries %ld (avg %.1f bonds)\n",L,carrymax,tailmax,batches,rounds_cta,rounds_tail,tails,carries,carries?carried_tot/(double)carries:0.);
  return 0;
}
