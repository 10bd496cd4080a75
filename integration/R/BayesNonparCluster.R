## Drop-in body for R/BayesNonparCluster.R of the reference: same signature, same return
## object; everything between the prior set-up (:21-151) and the packing of the result
## (:279-287) is one .Call into libclonalscope_b200.so.  `rng = "R"` (default here) makes the
## device consume base R's Mersenne-Twister stream exactly like the reference's loop, so the
## returned Zall is the one the pure-R body would have produced; `rng = "philox"` is the fast
## counter-based stream.
BayesNonparCluster=function(Xir=NULL,cna_states_WGS=NULL,alpha=2, beta=2, niter=200, sigmas0=NULL, U0=NULL, Z0=NULL, clust_mode='all', seed=200, verbose=FALSE,
                            rng=getOption("clonalscope.rng", "R"), device=getOption("clonalscope.device", -1L)){
  N=nrow(Xir); R=ncol(Xir)
  Xir=as.matrix(Xir); storage.mode(Xir)="double"
  cna_states_WGS=as.numeric(cna_states_WGS)
  ## priors: verbatim from the reference (:32-71)
  if(is.null(U0)){
    wU0=FALSE
    if(!is.null(cna_states_WGS)){ U0=matrix(c(rep(1,R), cna_states_WGS), byrow=T, ncol=R)
    }else{ U0=matrix(rep(1,R), byrow=T, ncol=R); cna_states_WGS=U0[1,] }
  }else if(ncol(U0)==R){
    wU0=TRUE
    U0=U0[1:max(which(!is.na(U0[,1]))),, drop=F]
    if(is.null(cna_states_WGS)| (length(cna_states_WGS)!=R)){
      message("Invalid/no cna_states_WGS detected. Use the one the same as that in the U0.")
      cna_states_WGS = if(nrow(U0)!=1) U0[which(!is.na(U0[,1]))[2],] else U0[1,]
    }
  }else stop("Please specify a U0 matrix with ncol the same as that of Xir!")
  if(!is.null(sigmas0)){
    if(length(sigmas0)==1) sigmas0=rep(sigmas0, R) else if(length(sigmas0)!=R) stop("Please specify sigmas0 with length R or length 1.")
  }else sigmas0=rep(0.25,R)
  storage.mode(U0)="double"
  ## Z0 (:91-112): argmax of the device log-likelihood matrix
  if(is.null(Z0)){
    PP=.Call("cs_R_loglik_matrix", Xir, U0, as.numeric(sigmas0), PACKAGE="Clonalscope")
    Z0=apply(PP, 1, which.max)
  }else if(length(Z0)!=N){ stop("Please specify a vector with length the same as the number of cells!")
  }else Z0[is.na(Z0)]=0
  colnames(U0)=names(sigmas0)=paste0("R", 1:R)

  set.seed(seed)                       # reference side effect (:156); see rng_state below
  res=.Call("cs_R_ncrp_gibbs", Xir, U0, as.integer(Z0), as.numeric(sigmas0), as.numeric(alpha), as.numeric(beta),
            as.integer(niter), as.integer(seed), as.integer(c(R=0L, philox=1L)[[rng]]), as.integer(device),
            PACKAGE="Clonalscope")
  if(rng=="R") assign(".Random.seed", c(.Random.seed[1], res$rng_state), envir=globalenv())  # leave R's stream where the loop would

  ## unpack into the reference's object (:238-287)
  Uall=vector("list",niter); sigma_all=vector("list", niter)
  u_rows=matrix(res$u_rows, ncol=R, byrow=TRUE)
  for(tt in 1:niter){
    Ut=matrix(NA_real_, nrow=res$kt[tt], ncol=R)
    idx=seq_len(res$u_off[tt+1]-res$u_off[tt])+res$u_off[tt]
    Ut[res$u_labels[idx],]=u_rows[idx,,drop=F]
    colnames(Ut)=paste0("R", 1:R); rownames(Ut)=paste0('Cluster',1:nrow(Ut))
    Uall[[tt]]=Ut
    s=res$sigma_all[,tt]; names(s)=paste0("R", 1:R); sigma_all[[tt]]=s
    if(verbose==TRUE){ print(paste0("Iteration:", tt)); print(paste0("nclusters=",nrow(Ut))); print(table(res$Zall[tt,]))
    }else if(tt%%20==0) print(paste0("Iteration:", tt))
  }
  Zall=res$Zall
  rownames(Zall)=paste0("Iter",1:niter); colnames(Zall)=rownames(Xir)
  names(cna_states_WGS)=colnames(Xir)
  if(length(unique(Z0))==1) U0=matrix(rep(1,R), byrow=T, ncol=R)       # :141-146
  results=list(Zall=Zall, Uall=Uall, sigma_all=sigma_all, Likelihood=res$Likelihood)
  priors=list(Z0=Z0, U0=U0, sigmas0=sigmas0, alpha=alpha, beta=beta, cna_states_WGS=cna_states_WGS, L0=res$L0, wU0=wU0, seed=seed)
  return(list(results=results, priors=priors, data=Xir))
}
