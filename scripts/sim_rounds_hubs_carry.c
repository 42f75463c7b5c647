// Carried-over bonds AND K hubs (env K, THR), with a check of every merge record against the sequential run.
//   gcc -O2 -o sim4 scripts/sim_rounds_hubs_carry.c;  K=2 THR=256 ./sim4 L carrymax tailmax [seed]
// This is the executable specification of the sweep design proposed for the next round.
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
  int L=atoi(argv[1]); int carrymax=atoi(argv[2]); int tailmax=atoi(argv[3]); int K=getenv("K")?atoi(getenv("K")):1; int thr=getenv("THR")?atoi(getenv("THR")):256;
  rng ^= (argc>4? strtoull(argv[4],0,10)*0x9E3779B97F4A7C15ull:0);
  int N=L*L, M=2*L*(L-1);
  int *eu=malloc(4*M),*ev=malloc(4*M); int m=0;
  for(int x=0;x<L;x++)for(int y=0;y<L;y++){ int id=x*L+y; if(y+1<L){eu[m]=id;ev[m]=id+1;m++;} if(x+1<L){eu[m]=id;ev[m]=id+L;m++;} }
  int *perm=malloc(4*M); for(int i=0;i<M;i++)perm[i]=i; for(int i=M-1;i>0;i--){int j=xr()%(i+1);int t=perm[i];perm[i]=perm[j];perm[j]=t;}
  par=malloc(4*N); sz=malloc(4*N); for(int i=0;i<N;i++){par[i]=i;sz[i]=1;}
  long long *rec_seq=calloc(M,8), *rec_sim=calloc(M,8);
  for(int n=0;n<M;n++){ int a=find(eu[perm[n]]), b=find(ev[perm[n]]); if(a==b)continue; int lo=sz[a]<sz[b]?sz[a]:sz[b], hi=sz[a]<sz[b]?sz[b]:sz[a]; rec_seq[n]=((long long)lo<<32)|hi; if(sz[a]<sz[b]){int t=a;a=b;b=t;} par[b]=a; sz[a]+=sz[b]; }
  for(int i=0;i<N;i++){par[i]=i;sz[i]=1;}
  static int bidx[T+CMAX];
  int *owner=malloc(4*N); for(int i=0;i<N;i++)owner[i]=1<<30;
  long rounds_cta=0, rounds_tail=0, tails=0, batches=0, carried_tot=0, carries=0;
  int hubs[8]; int nh=0;
  int ru[T+CMAX],rv[T+CMAX],pend[T+CMAX]; int ncar=0;
  // verification of sequential semantics: record (size a,size b) per bond and compare with sequential run
  for(int n0=0;n0<M || ncar>0;n0+=T){
    int cnt = n0<M ? (M-n0<T?M-n0:T) : 0; int tot=ncar+cnt;
    for(int i=0;i<cnt;i++){ int e=perm[n0+i]; ru[ncar+i]=find(eu[e]); rv[ncar+i]=find(ev[e]); pend[ncar+i]=ru[ncar+i]!=rv[ncar+i]; bidx[ncar+i]=n0+i; }
    int last = n0+T>=M;
    int intail=0;
    for(;;){
      int left=0; for(int i=0;i<tot;i++){ if(pend[i]){ ru[i]=find(ru[i]); rv[i]=find(rv[i]); pend[i]=ru[i]!=rv[i]; } left+=pend[i]; }
      if(!left)break;
      if(!last && left<=carrymax) break;       // carry them
      if(left<=tailmax && !intail){ intail=1; tails++; }
      // hubs: the K largest clusters among the previous hubs and the roots touched by the batch
      int cand[2*(T+CMAX)+8]; int nc=0; for(int h=0;h<nh;h++)cand[nc++]=find(hubs[h]); for(int i=0;i<tot;i++){cand[nc++]=ru[i];cand[nc++]=rv[i];}
      nh=0; for(int k=0;k<K;k++){ int best=-1; for(int c=0;c<nc;c++){ int x=cand[c]; int dup=0; for(int h=0;h<nh;h++) if(hubs[h]==x)dup=1; if(dup)continue; if(best<0||sz[x]>sz[best]||(sz[x]==sz[best]&&x>best))best=x; } if(best>=0 && (k==0 || sz[best]>=thr))hubs[nh++]=best; }
      int star[T+CMAX], o[T+CMAX], hb[T+CMAX];
      for(int i=0;i<tot;i++){ star[i]=0; if(!pend[i])continue; int hu=-1,hv=-1; for(int h=0;h<nh;h++){ if(ru[i]==hubs[h])hu=h; if(rv[i]==hubs[h])hv=h; }
        if(hu>=0&&hv>=0) star[i]=0; else if(hu>=0){star[i]=1;o[i]=rv[i];hb[i]=hu;} else if(hv>=0){star[i]=1;o[i]=ru[i];hb[i]=hv;} }
      for(int i=0;i<tot;i++){ if(!pend[i])continue; if(star[i]){ if(owner[o[i]]>i)owner[o[i]]=i; } else { if(owner[ru[i]]>i)owner[ru[i]]=i; if(owner[rv[i]]>i)owner[rv[i]]=i; } }
      int own[T+CMAX]; int bmin=1<<30;
      int firststar[8]; for(int h=0;h<8;h++)firststar[h]=1<<30;
      for(int i=0;i<tot;i++) if(pend[i]&&star[i]&&firststar[hb[i]]>i) firststar[hb[i]]=i;
      for(int i=0;i<tot;i++){ own[i]=0; if(!pend[i])continue; if(star[i]) own[i]= owner[o[i]]==i && owner[hubs[hb[i]]]>i; else { own[i]= owner[ru[i]]==i && owner[rv[i]]==i; if(own[i]) for(int h=0;h<nh;h++) if((ru[i]==hubs[h]||rv[i]==hubs[h]) && firststar[h]<i) own[i]=0; } if(!own[i] && bmin>i) bmin=i; }
      int merged=0;
      for(int i=0;i<tot;i++){ if(!pend[i]||!own[i]||star[i])continue; int a=ru[i],b=rv[i]; if(sz[a]<sz[b]){int t=a;a=b;b=t;} rec_sim[bidx[i]]=((long long)sz[b]<<32)|sz[a]; par[b]=a; sz[a]+=sz[b]; pend[i]=0; merged++; }
      for(int i=0;i<tot;i++){ if(!pend[i]||!own[i]||!star[i])continue; if(i>=bmin)continue; int a=find(hubs[hb[i]]),b=o[i]; if(a==b){pend[i]=0;continue;} if(sz[a]<sz[b]){int t=a;a=b;b=t;} rec_sim[bidx[i]]=((long long)sz[b]<<32)|sz[a]; par[b]=a; sz[a]+=sz[b]; pend[i]=0; merged++; }
      for(int i=0;i<tot;i++){ owner[ru[i]]=1<<30; owner[rv[i]]=1<<30; } for(int h=0;h<nh;h++) owner[hubs[h]]=1<<30;
      if(!merged){ fprintf(stderr,"stuck\n"); return 1; }
      if(intail)rounds_tail++; else rounds_cta++;
    }
    // compact the pending bonds (in order) to the front
    int k=0; for(int i=0;i<tot;i++) if(pend[i]){ ru[k]=ru[i]; rv[k]=rv[i]; bidx[k]=bidx[i]; pend[k]=1; k++; }
    ncar=k; if(k){carries++; carried_tot+=k;}
    batches++;
    if(n0>=M && ncar==0)break;
  }
  { long bad=0; for(int n=0;n<M;n++) if(rec_seq[n]!=rec_sim[n]) bad++; printf("records differing from the sequential run: %ld\n",bad); }
  printf("L=%d carrymax=%d tailmax=%d: batches %ld, CTA rounds %ld, tail rounds %ld (tails %ld), carries %ld (avg %.1f bonds)\n",L,carrymax,tailmax,batches,rounds_cta,rounds_tail,tails,carries,carries?carried_tot/(double)carries:0.);
  return 0;
}
